#!/usr/bin/env python
"""bench.py -- images/sec of the YOLOv3 detect hot path at 608x608, batch 32 per GPU (BASELINE.json configs[2]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg3|cfg1|cfg2|cfg4|cfg5]

A step = one pass of the hot path over one batch of synthetic images: 75 fused convolutions (tcgen05 tensor cores, fp16
in / fp32 accumulate) -> 3-scale anchor decode -> box filter / sort / IOU / greedy NMS (yb_detect), i.e. what test.py:35-36
of the reference computes per batch (net(imgs) + postprocessing(cat(det), conf 0.5, nms 0.4)).  Weights are random-init
(BN-calibrated, seeded) of the real architecture; images are uniform noise -- there is no network for datasets.

Printed JSON (one line, rank 0):
  value         whole-job images/sec with the input batches already resident in HBM (CUDA events on the launch stream,
                max over ranks), K timed steps after W warm-up steps;
  e2e           the same through the public Python API with HOST buffers -- the reference caller's contract, an fp32
                [B,3,608,608] batch (test.py:32): every step copies its batch from pinned host memory and reads the
                detections back (copy engine overlapped with compute by double buffering; both inside the timed region);
  e2e_u8_frames / e2e_f16_input   the same from uint8 camera frames (letterboxed on the device, yb_letterbox) and from
                fp16 images read by the stem directly: 1/5 and 1/2 of the PCIe bytes, at every N;
  roofline      the convolution stack: algorithmic 2*MAC FLOPs / CUDA-event time of the conv section of the timed steps,
                against the BURST bf16 peak of MEASURED_PEAKS.json (the K-step region lasts ~0.1 s); `sustained` = the
                same over >= 300 back-to-back steps (the board reaches its power cap) against the sustained peak;
  parity        computed outside the timed region: how the benched fp16 path's final detections compare with the fp32
                CPU oracle's on images of the bench batch (matched boxes, set IOU), and the fp32-grade mode's deviation;
  parity_mode   images/sec of precision='fp32' (YB_MODE_FP32_TC: fp32-grade parity with the reference on the tensor
                cores) on the same workload, with its own roofline;
  other_configs BASELINE configs 1, 2 and 4 (416 single image latency, backbone 256 b64, NMS stress), short runs;
  multi_gpu_check (N > 1) the weight broadcast overwrote deliberately different weights and the gathered detections equal
                every rank's local ones;
  cpu_baseline  the reference's own modules (oracle/_ref: darknet.YoloNet + utils.postprocessing, unmodified) on the host
                cores, same protocol as --impl reference: steps of 4 images.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONF_THR, NMS_THR = 0.5, 0.4
METRIC = "images/sec at 608x608 batch-32 (YOLOv3 detect: backbone + 3-scale decode + NMS)"
REF_IMAGES_PER_STEP = 4          # one protocol for both CPU legs: steps of 4 images of the same workload
TRAFFIC_CSV = os.path.join("profiles", "r04_conv_metrics.csv")


def ncu_conv_traffic():
    """DRAM bytes (read+write) of the 74 convolution launches of one step, from the committed ncu capture of the CURRENT
    kernel set (608x608 batch 32 only)."""
    import csv
    p = os.path.join(ROOT, TRAFFIC_CSV)
    if not os.path.exists(p):
        return None
    rows = list(csv.reader(open(p)))
    hdr = next((r for r in rows if "Kernel Name" in r), None)
    if hdr is None:
        return None
    i_name, i_val, i_unit = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    total = 0.0
    for r in rows:
        if len(r) > i_unit and r[i_name] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            total += float(r[i_val].replace(",", "")) * scale.get(r[i_unit], 1.0)
    return total or None


def _flatten(d, prefix=""):
    if isinstance(d, dict):
        for k, v in d.items():
            yield from _flatten(v, f"{prefix}.{k}".lower() if prefix else str(k).lower())
    elif isinstance(d, (list, tuple)):
        for i, v in enumerate(d):
            yield from _flatten(v, f"{prefix}[{i}]")
    elif isinstance(d, (int, float)) and not isinstance(d, bool):
        yield prefix, float(d)


def parse_peaks(d):
    """(sustained tensor TFLOP/s, HBM GB/s, burst tensor TFLOP/s) out of the driver-written MEASURED_PEAKS.json, whose exact
    key names this repository does not control: known names first, then any numeric leaf whose path says what it is.
    Any of the three may be None."""
    flat = list(_flatten(d))
    by = dict(flat)
    tensor = by.get("bf16_tflops_sustained")
    burst = by.get("bf16_tflops_burst", by.get("bf16_tflops") if tensor is not None else None)
    hbm = by.get("hbm_gbs")
    cand = [(k, v) for k, v in flat if any(t in k for t in ("bf16", "tflop", "tf/s", "tensor")) and 100.0 <= v <= 5000.0]
    if tensor is None:
        sus = [v for k, v in cand if "sustain" in k]
        other = [v for k, v in cand if "burst" not in k and "peak" not in k]
        tensor = sus[0] if sus else (other[0] if other else (min(v for _, v in cand) if cand else None))
    if burst is None:
        b = [v for k, v in cand if "sustain" not in k]
        burst = max(b) if b else tensor
    if hbm is None:
        c2 = [v for k, v in flat if any(t in k for t in ("hbm", "dram", "copy", "gb/s", "gbs", "gbps", "bandwidth")) and 500.0 <= v <= 20000.0]
        hbm = c2[0] if c2 else None
    return tensor, hbm, burst


def peaks():
    """Roofline denominators: MEASURED_PEAKS.json when the driver has written it, else the profiling recipe's fallback."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    tensor = hbm = burst = None
    if os.path.exists(p):
        try:
            tensor, hbm, burst = parse_peaks(json.load(open(p)))
        except Exception as e:  # noqa: BLE001 -- a malformed file must not take the bench down
            print(f"[bench] MEASURED_PEAKS.json unreadable ({e}); using the fallback peaks", file=sys.stderr)
    if tensor is None and hbm is None:
        return dict(tensor=1400.0, burst=1400.0, hbm=6650.0, src="fallback")
    src = "measured" if tensor is not None and hbm is not None else "measured+fallback"
    tensor = tensor if tensor is not None else 1400.0
    return dict(tensor=tensor, burst=burst if burst is not None else tensor, hbm=hbm if hbm is not None else 6650.0, src=src)


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / power / throttle reasons while a timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([v.strip() for v in out.split(",")])
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.05)

    def finish(self):
        self.stop_flag = True
        self.join(timeout=3)
        return self.summary()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i].lower().startswith("active") for s in self.samples)]
        pw = [float(s[6]) for s in self.samples if len(s) > 6 and s[6].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference itself (oracle/_ref), or -- where that directory did not travel -- the oracle port
# ---------------------------------------------------------------------------------------------------------------------
def make_cpu_step(size, images):
    """Returns (step, kind, description): step() runs forward + decode + postprocessing on `images` images on the host."""
    import torch
    from oracle import ref_loader
    from yolo_v3_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)            # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every core
    sd = synth.make_state_dict(seed=1234, recipe="calibrated")
    x = synth.make_images(images, size, size, seed=0)
    if ref_loader.available():
        darknet, utils = ref_loader.load()
        net = darknet.YoloNet((size, size))
        net.load_state_dict(sd)
        net.eval()

        def step():
            with ref_loader.cpu_only(), torch.no_grad():
                det = torch.cat(net(x, None), 1)           # test.py:35-36
                return utils.postprocessing(det, 80, CONF_THR, NMS_THR)
        return step, "reference", (f"the reference's own darknet.YoloNet.forward + utils.postprocessing (oracle/_ref, unmodified; "
                                   f"Tensor.cuda shimmed to identity so that the path stays on the host), torch {torch.__version__} fp32")
    from oracle import yolo_oracle as O

    def step():
        return O.postprocessing(torch.cat(O.forward(sd, x), 1), 80, CONF_THR, NMS_THR)
    return step, "port", f"oracle/yolo_oracle.py (restates darknet.py / yololayer.py / utils.py; oracle/_ref absent), torch {torch.__version__} fp32"


def time_cpu(size, images, steps, warmup):
    import torch
    step, kind, desc = make_cpu_step(size, images)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return dict(value=images * steps / dt, unit="images/sec", cores=torch.get_num_threads(), kind=kind,
                sample=f"{steps} steps x {images} images {size}x{size} in {dt:.1f}s after {warmup} warm-up steps: {desc}"), dt


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the box's host cores, W warm-up + exactly K
    steps of REF_IMAGES_PER_STEP images (a bounded sample of the batch-32 workload); rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    base, dt = time_cpu(args.size, REF_IMAGES_PER_STEP, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "images/sec", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"yolov3_{args.size}x{args.size}_b{args.batch}_detect", "conf_thr": CONF_THR, "nms_thr": NMS_THR,
                       "note": f"CPU path; each step is a bounded sample of {REF_IMAGES_PER_STEP} images of the batch-{args.batch} workload"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ---------------------------------------------------------------------------------------------------------------------
# helpers of the GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def match_detections(got, ref, iou_thr=0.5):
    """Greedy one-to-one matching of two detection lists of one image (rows x1,y1,x2,y2,obj,score,cls): same class and
    IOU > iou_thr.  Returns (matched, len(got), len(ref), mean IOU of the matches, max |score difference| of the matches)."""
    import torch
    if len(got) == 0 or len(ref) == 0:
        return 0, len(got), len(ref), 0.0, 0.0
    g, r = got.float(), ref.float()
    x1 = torch.max(g[:, None, 0], r[None, :, 0]); y1 = torch.max(g[:, None, 1], r[None, :, 1])
    x2 = torch.min(g[:, None, 2], r[None, :, 2]); y2 = torch.min(g[:, None, 3], r[None, :, 3])
    inter = (x2 - x1).clamp(min=0) * (y2 - y1).clamp(min=0)
    ag = ((g[:, 2] - g[:, 0]) * (g[:, 3] - g[:, 1]))[:, None]
    ar = ((r[:, 2] - r[:, 0]) * (r[:, 3] - r[:, 1]))[None, :]
    iou = inter / (ag + ar - inter).clamp(min=1e-9)
    iou = torch.where(g[:, None, 6] == r[None, :, 6], iou, torch.zeros_like(iou))
    matched, ious, dscore = 0, [], 0.0
    used = torch.zeros(len(r), dtype=torch.bool)
    for i in torch.argsort(iou.max(1).values, descending=True).tolist():
        row = iou[i].clone()
        row[used] = 0
        j = int(row.argmax())
        if float(row[j]) > iou_thr:
            used[j] = True
            matched += 1
            ious.append(float(row[j]))
            dscore = max(dscore, abs(float(g[i, 5] - r[j, 5])))
    return matched, len(g), len(r), (sum(ious) / len(ious) if ious else 0.0), dscore


def parity_report(net16, sd, x_host, size):
    """A.5 L2b of the survey: final detections of the benched fp16 path vs the fp32 CPU oracle on the same images at the
    bench thresholds, plus the fp32-grade tensor-core mode's deviation.  Outside every timed region."""
    import torch
    from oracle import yolo_oracle as O
    from yolo_v3_b200 import YoloNet
    n = x_host.shape[0]
    with torch.no_grad():
        ref_logits = O.head_logits(sd, x_host)
        ref_det = torch.cat(O.forward(sd, x_host), 1)
    ref = O.postprocessing(ref_det.clone(), 80, CONF_THR, NMS_THR)
    xd = x_host.cuda()
    out = {"images": n, "conf_thr": CONF_THR, "nms_thr": NMS_THR, "oracle": "oracle/yolo_oracle.py (fp32, CPU), pinned to the reference's outputs"}

    def compare(dets):
        tot_m = tot_g = tot_r = 0
        ious, ds = [], 0.0
        for g, r in zip(dets if dets else [torch.zeros(0, 7)] * n, ref if ref else [torch.zeros(0, 7)] * n):
            m, ng, nr, mi, d = match_detections(g.cpu(), r)
            tot_m += m; tot_g += ng; tot_r += nr
            if m:
                ious.append(mi)
            ds = max(ds, d)
        return {"detections": tot_g, "oracle_detections": tot_r, "matched": tot_m, "only_ours": tot_g - tot_m,
                "only_oracle": tot_r - tot_m, "set_iou": tot_m / max(1, tot_g + tot_r - tot_m),
                "mean_box_iou_of_matches": sum(ious) / len(ious) if ious else None, "max_score_diff_of_matches": ds}

    mx = max(float(l.abs().max()) for l in ref_logits)
    det16 = torch.cat(net16(xd, None), 1).cpu()
    d16 = (det16 - ref_det).abs()
    l16 = [l.cpu() for l in net16.head_logits(xd)]
    out["fp16"] = dict(compare(net16.detect(xd, CONF_THR, NMS_THR)),
                       logits_max_abs_diff_over_max_logit=max(float((a - b).abs().max()) for a, b in zip(l16, ref_logits)) / mx,
                       max_abs_diff_xy_px=float(d16[..., :2].max()), max_abs_diff_obj_cls=float(d16[..., 4:].max()))
    net32 = YoloNet((size, size), precision="fp32")
    net32.load_state_dict(sd)
    net32 = net32.cuda().eval()
    det32 = torch.cat(net32(xd, None), 1).cpu()
    d32 = (det32 - ref_det).abs()
    l32 = [l.cpu() for l in net32.head_logits(xd)]
    out["fp32_tc"] = dict(compare(net32.detect(xd, CONF_THR, NMS_THR)),
                          logits_max_abs_diff_over_max_logit=max(float((a - b).abs().max()) for a, b in zip(l32, ref_logits)) / mx,
                          max_abs_diff_xy_px=float(d32[..., :2].max()), max_abs_diff_obj_cls=float(d32[..., 4:].max()),
                          tolerance="north star: 1e-4 on scores, logits within 1e-4 * max|logit|")
    del net32
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import ctypes
    import numpy as np
    import torch
    import torch.distributed as dist
    from yolo_v3_b200 import YoloNet, synth, topology
    from yolo_v3_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = args.config
    B, S = args.batch, args.size
    if cfg == "cfg5":                      # strong scaling: global batch 256 split over the ranks
        B = 256 // world
    N = topology.num_boxes(S, S)
    pk = peaks()
    stream = torch.cuda.current_stream()
    lib = _lib.load()
    cap = 512

    sd = synth.make_state_dict(seed=1234, recipe="calibrated")

    def make_net(precision, size=S, state=None):
        n = YoloNet((size, size), precision=precision)
        n.load_state_dict(state if state is not None else sd)
        return n.cuda().eval()

    # ---- multi-GPU set-up: ranks > 0 start from DIFFERENT weights, so the broadcast below has something to overwrite ----
    mg = None
    if world > 1 and rank != 0:
        # (conv weights scaled; the stem's BN bias shifted too: the first kernel takes the stem's scale / bias as kernel
        # parameters from a host mirror, which the broadcast must refresh)
        sd_other = {k: (v * (1.0 + 0.05 * rank) if k.endswith("conv.weight") else
                        v + 0.02 * rank if k == "feature.mlist.0.bn.bias" else v.clone()) for k, v in sd.items()}
        net = make_net(args.precision, S, sd_other)
    else:
        net = make_net(args.precision)
    xs = [synth.make_images(B, S, S, seed=100 * rank + i).cuda() for i in range(2)]     # two resident batches > L2, alternated
    comm = None
    if world > 1:
        from yolo_v3_b200 import parallel
        comm = parallel.DetectionGather(net, rank, world, B, cap)
        shared = synth.make_images(4, S, S, seed=4242).cuda()                            # the same images on every rank

        def gather_all(t):
            out = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(out, t.contiguous())
            return out

        rows_b, counts_b, _, _ = net.detect_raw(shared, CONF_THR, NMS_THR, False, True, cap)   # creates + finalises the engine
        before = gather_all(rows_b.clone())
        before_c = gather_all(counts_b.clone())
        comm.broadcast_weights()
        rows_a, counts_a, _, _ = net.detect_raw(shared, CONF_THR, NMS_THR, False, True, cap)
        after = gather_all(rows_a.clone())
        after_c = gather_all(counts_a.clone())

        def same(r0, c0, r1, c1):
            if not torch.equal(c0, c1):
                return False
            return all(torch.equal(r0[i, :int(c0[i])], r1[i, :int(c0[i])]) for i in range(r0.shape[0]))
        differed_before = all(not same(before[0], before_c[0], before[r], before_c[r]) for r in range(1, world))
        equal_after = all(same(after[0], after_c[0], after[r], after_c[r]) for r in range(1, world))
        mg = {"weights_differed_before_broadcast": bool(differed_before), "detections_equal_after_broadcast": bool(equal_after)}

    def step(i):
        rows, counts, src, cand = net.detect_raw(xs[i & 1], CONF_THR, NMS_THR, False, True, cap)
        if comm is not None:
            return comm.allgather(rows, counts)
        return rows, counts

    if comm is not None:
        # the gathered tensor must hold every rank's local rows in rank-major order (checked through torch's own NCCL path)
        rows_l, counts_l, _, _ = net.detect_raw(xs[0], CONF_THR, NMS_THR, False, True, cap)
        ar, ac = comm.allgather(rows_l, counts_l)
        ar, ac = ar.clone(), ac.clone()
        lr, lc = gather_all(rows_l.clone()), gather_all(counts_l.clone())
        ok = True
        for r in range(world):
            ok = ok and torch.equal(ac[r * B:(r + 1) * B], lc[r])
            for i in range(B):
                k = int(lc[r][i])
                ok = ok and torch.equal(ar[r * B + i, :k], lr[r][i, :k])
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        mg["gathered_rows_equal_local_rows"] = bool(int(flag.item()))
        mg["status"] = "ok" if all(v for v in mg.values() if isinstance(v, bool)) else "FAILED"

    def timed_steps(n_steps, warmup, sample_clocks):
        """W warm-up + exactly n_steps steps between barrier + synchronize on both sides; max over ranks; CUDA events."""
        for i in range(warmup):
            step(i)
        torch.cuda.synchronize()
        sampler = ClockSampler(local) if (sample_clocks and rank == 0) else None
        if sampler:
            sampler.start()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = lib.yb_launch_count(net._ctx)
        e0.record(stream)
        out = None
        for i in range(n_steps):
            out = step(i)
        e1.record(stream)
        torch.cuda.synchronize()
        launched = lib.yb_launch_count(net._ctx) - n0     # kernels of libyolo_b200.so launched inside the timed region
        if world > 1:
            dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), out, (sampler.finish() if sampler else None), int(launched)

    step(0)
    torch.cuda.synchronize()
    ctx = net._ctx
    net.freeze_weights()                                   # weights are fixed from here on: no per-call change detection
    ms, out, clocks, launches = timed_steps(args.steps, args.warmup, True)
    counts_h = out[1].cpu()
    assert int(counts_h.max()) <= cap, "detection capacity overflow in the timed region"

    # ---- sustained leg: >= 300 steps back to back (the board reaches its power cap), clocks + power recorded ----
    sus_steps = max(args.sustained, 0)
    sus = None
    if sus_steps:
        ms_s, _, clocks_s, _ = timed_steps(sus_steps, 3, True)
        sus = {"steps": sus_steps, "ms_per_step": ms_s / sus_steps, "value": B * world * sus_steps / (ms_s * 1e-3), "clocks": clocks_s}

    # ---- end to end: pinned host batches in, detections out, every step ----
    copy_stream = torch.cuda.Stream()
    ready = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]
    h_rows = torch.empty(B * (world if comm else 1), cap, 7).pin_memory()
    h_counts = torch.empty(B * (world if comm else 1), dtype=torch.int32).pin_memory()
    d2h = int(h_rows.numel() * 4 + h_counts.numel() * 4)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def e2e_loop(n, host, devb, to_input):
        """Double-buffered: the H2D copy of step i+1 (copy stream) overlaps the compute of step i; detections of every
        step go back to pinned host memory.  to_input maps the device copy of the host batch to the network input."""
        with torch.cuda.stream(copy_stream):
            devb[0].copy_(host[0], non_blocking=True)
            ready[0].record(copy_stream)
        for i in range(n):
            cur, nxt = i & 1, (i + 1) & 1
            if i + 1 < n:
                with torch.cuda.stream(copy_stream):
                    if i >= 1:
                        copy_stream.wait_event(free[nxt])
                    devb[nxt].copy_(host[nxt], non_blocking=True)
                    ready[nxt].record(copy_stream)
            stream.wait_event(ready[cur])
            rows, counts, _, _ = net.detect_raw(to_input(devb[cur]), CONF_THR, NMS_THR, False, True, cap)
            free[cur].record(stream)
            if comm is not None:
                rows, counts = comm.allgather(rows, counts)
            h_rows.copy_(rows, non_blocking=True)
            h_counts.copy_(counts, non_blocking=True)
        stream.synchronize()

    def time_e2e(host, devb, to_input):
        e2e_loop(max(3, args.warmup), host, devb, to_input)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        e0.record(stream)
        e2e_loop(args.steps, host, devb, to_input)
        e1.record(stream)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        v = max(e0.elapsed_time(e1), wall)       # the first H2D precedes e0 on the compute stream: take the longer clock
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def e2e_entry(t_ms, h2d, what):
        return {"value": B * world * args.steps / (t_ms * 1e-3), "unit": "images/sec", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": t_ms / args.steps, "h2d_GBps_per_rank": h2d / (t_ms / args.steps * 1e-3) / 1e9, "input": what}

    hx = [synth.make_images(B, S, S, seed=7 + i).pin_memory() for i in range(2)]
    dx = [torch.empty_like(xs[0]) for _ in range(2)]
    e2e = e2e_entry(time_e2e(hx, dx, lambda x: x), B * 3 * S * S * 4,
                    f"fp32 [B,3,{S},{S}] in [0,1], what the reference's predict() moves with .cuda() (test.py:32)")
    extra = {}
    FH, FW = 480, 640
    try:                                                    # (a) camera frames: uint8 in, letterbox on the device (the N1 row)
        from yolo_v3_b200.utils import letterbox_batch
        frames = np.stack([synth.make_photo(FH, FW, 90 + i) for i in range(min(B, 8))])
        frames = np.concatenate([frames] * ((B + len(frames) - 1) // len(frames)))[:B]
        hu = [torch.from_numpy(frames).pin_memory(), torch.from_numpy(np.ascontiguousarray(frames[::-1])).pin_memory()]
        du = [torch.empty(B, FH, FW, 3, dtype=torch.uint8, device=dev) for _ in range(2)]
        extra["e2e_u8_frames"] = e2e_entry(time_e2e(hu, du, lambda u: letterbox_batch(list(u), (S, S))[0]), B * FH * FW * 3,
                                           f"uint8 [B,{FH},{FW},3] frames, letterboxed to {S}x{S} on the device (yb_letterbox) inside the timed region")
        del hu, du
    except Exception as e:  # noqa: BLE001 -- an optional leg must not take the contract line down
        extra["e2e_u8_frames"] = {"error": f"{type(e).__name__}: {e}"}
        torch.cuda.synchronize()
    if args.precision == "fp16":                            # (b) fp16 images read by the stem directly: bit-identical detections
        try:
            hh = [h.half().pin_memory() for h in hx]
            dh = [torch.empty(B, 3, S, S, dtype=torch.float16, device=dev) for _ in range(2)]
            extra["e2e_f16_input"] = e2e_entry(time_e2e(hh, dh, lambda x: x), B * 3 * S * S * 2,
                                               f"fp16 [B,3,{S},{S}] read by the stem directly (yb_set_input_dtype; bit-identical detections)")
            del hh, dh
        except Exception as e:  # noqa: BLE001
            extra["e2e_f16_input"] = {"error": f"{type(e).__name__}: {e}"}
            torch.cuda.synchronize()
    del hx, dx

    # ---- roofline: section times of the same step, CUDA events per section on the launch stream, recorded without
    # synchronising so that the steps still run back to back as in the timed loop; one query at the end averages them ----
    _lib.check(lib.yb_set_profiling(ctx, 3), ctx)
    a, b, c = ctypes.c_float(), ctypes.c_float(), ctypes.c_float()
    for i in range(3):
        net.detect_raw(xs[i & 1], CONF_THR, NMS_THR, False, True, cap)
    lib.yb_get_section_ms(ctx, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))      # discard the warm-up sets
    for i in range(min(args.steps, 16)):
        net.detect_raw(xs[i & 1], CONF_THR, NMS_THR, False, True, cap)
    lib.yb_get_section_ms(ctx, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
    conv_ms, dec_ms, post_ms = a.value, b.value, c.value
    conv_ms_sus = None
    if sus_steps:                                           # conv share inside a long run: sample the last 16 of 100 more steps
        for i in range(100):
            net.detect_raw(xs[i & 1], CONF_THR, NMS_THR, False, True, cap)
        lib.yb_get_section_ms(ctx, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
        conv_ms_sus = a.value
    layer_ms = []
    if args.layers:
        _lib.check(lib.yb_set_profiling(ctx, 2), ctx)      # + one event per launch of the conv section
        for i in range(3):
            net.detect_raw(xs[i & 1], CONF_THR, NMS_THR, False, True, cap)
            buf = (ctypes.c_float * 96)()
            n = lib.yb_get_layer_ms(ctx, buf, 96)
            cur = [buf[j] for j in range(min(n, 96))]
            layer_ms = cur if not layer_ms else [x + y for x, y in zip(layer_ms, cur)]
        layer_ms = [v / 3 for v in layer_ms]
    _lib.check(lib.yb_set_profiling(ctx, 0), ctx)

    if rank != 0:
        # every collective of the run is behind us: the other ranks leave, rank 0 finishes its single-GPU extras alone (it
        # keeps its engine -- and with it the library's NCCL communicator -- alive until the process exits: destroying a
        # communicator while the peers are gone or blocked elsewhere can wait for them for ever)
        torch.cuda.synchronize()
        dist.destroy_process_group()
        return
    torch.set_num_threads(os.cpu_count() or 1)             # torchrun exports OMP_NUM_THREADS=1; the oracle legs below use every core

    # ================= rank 0 only from here: kernels timed alone, parity, other configs, CPU baseline =================
    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record(stream)
        for _ in range(reps):
            fn()
        b_.record(stream)
        torch.cuda.synchronize()
        return a_.elapsed_time(b_) / reps

    def gbps(nbytes, t_ms):
        return {"ms": t_ms, "achieved_GBps": nbytes / (t_ms * 1e-3) / 1e9 if t_ms else None,
                "frac": nbytes / (t_ms * 1e-3) / 1e9 / pk["hbm"] if t_ms else None}

    from yolo_v3_b200.utils import letterbox_batch, postprocessing_raw
    _lib.check(lib.yb_set_profiling(ctx, 1), ctx)
    d_alone = 0.0
    for i in range(5):                                      # yb_forward: conv stack + standalone decode (writes det)
        dets = net(xs[i & 1], None)
        lib.yb_get_section_ms(ctx, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
        d_alone += b.value / 5
    _lib.check(lib.yb_set_profiling(ctx, 0), ctx)
    det_cat = net._last_det                                 # the concatenated [B,N,85] tensor det1..3 are views of
    p_alone = timed(lambda: postprocessing_raw(det_cat, 80, CONF_THR, NMS_THR, False, True, cap))
    photos = [torch.from_numpy(synth.make_photo(480, 640, 50 + i)).cuda() for i in range(B)]
    l_alone = timed(lambda: letterbox_batch(photos, (S, S)))
    lb_bytes = B * (480 * 640 * 3 + S * S * 12)
    del det_cat, dets, photos
    row_bytes = B * N * 85 * 4              # one pass over the [B,N,85] fp32 tensor (padded logit pitch ignored)

    flops = topology.conv_flops(S, S) * B
    achieved = flops / (conv_ms * 1e-3) / 1e12
    short_region = ms < 1000.0              # a sub-second timed region runs at burst clocks: compare with the burst peak
    peak = pk["burst"] if short_region else pk["tensor"]
    roof = {"bound": "tensor", "kernel": "stem_block_kernel (stem + layer 1) + conv_tc_kernel + conv_halo_kernel: 74 launches/step", "achieved": achieved,
            "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "peak_source": f"{pk['src']} {'burst' if short_region else 'sustained'} bf16 (MEASURED_PEAKS.json): the timed region lasts {ms:.0f} ms",
            "traffic": ncu_conv_traffic() if (S == 608 and B == 32 and args.precision == "fp16") else None,
            "traffic_note": f"DRAM read+write bytes of the conv launches of one step (ncu, {TRAFFIC_CSV}); algorithmic activation+weight bytes: 13.0e9 with every tensor once (11.5e9 without the fused stem output)",
            "conv_ms_per_step": conv_ms, "flops_per_step": flops}
    if sus is not None and conv_ms_sus:
        ach_s = flops / (conv_ms_sus * 1e-3) / 1e12
        roof["sustained"] = {"steps": sus["steps"], "ms_per_step": sus["ms_per_step"], "value": sus["value"], "conv_ms_per_step": conv_ms_sus,
                             "achieved": ach_s, "peak": pk["tensor"], "frac": ach_s / pk["tensor"], "clocks": sus["clocks"],
                             "note": "back-to-back steps; the board's power cap, not the kernel, sets the SM clock here"}

    line = {
        "metric": METRIC, "value": B * world * args.steps / (ms * 1e-3), "unit": "images/sec", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if cfg == "cfg5" else "weak",
        "vs_baseline": None, "dtype": "f16" if args.precision == "fp16" else "f32", "data": "synthetic",
        "config": {"workload": f"yolov3_{S}x{S}_b{B}_detect" + (" (BASELINE cfg5: global batch 256 sharded)" if cfg == "cfg5" else ""),
                   "batch_per_gpu": B, "global_batch": B * world, "img": S,
                   "conf_thr": CONF_THR, "nms_thr": NMS_THR, "precision": args.precision,
                   "l2": "inputs larger than L2 (two alternating 142 MB batches; activations 0.8 GB per layer)",
                   "parallelism": f"batch-sharded x{world}, weights broadcast once, detections all-gathered per step"},
        "e2e": e2e,
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roof,
        "roofline_hbm": {
            # inside the timed step (yb_detect): decode + score fused, reads the head maps once, det is never written
            "decode_score_fused": dict(gbps(row_bytes, dec_ms), bytes="logits in (upper bound: the objectness probe skips most cells)"),
            "post_scan_sort_nms_emit_ms": post_ms,
            # the same kernels behind the reference's separate calls, timed alone
            "decode": dict(gbps(2 * row_bytes, d_alone), bytes="logits in + det out"),
            "postprocess": dict(gbps(row_bytes, p_alone), bytes="det in"),
            "letterbox": dict(gbps(lb_bytes, l_alone), bytes="uint8 480x640 sources in + fp32 CHW canvases out"),
            "peak_GBps": pk["hbm"]},
        "detections_last_step": int(counts_h.sum()),
    }
    line.update(extra)
    if mg is not None:
        line["multi_gpu_check"] = mg["status"]
        line["multi_gpu_check_detail"] = mg

    # ---- everything below builds other engines: free this one first ----
    x_par = xs[0][:args.parity_images].cpu()
    if not args.quick:
        try:
            line["parity"] = parity_report(net, sd, x_par, S) if args.precision == "fp16" else None
        except Exception as e:  # noqa: BLE001
            line["parity"] = {"error": f"{type(e).__name__}: {e}"}
            torch.cuda.synchronize()
    if world == 1:
        del net
        torch.cuda.empty_cache()

    def simple_bench(n, batch_x, fn, steps):
        """W warm-up + `steps` calls of fn(i) on resident inputs, CUDA events; returns ms per step."""
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record(stream)
        for i in range(steps):
            fn(i)
        b_.record(stream)
        torch.cuda.synchronize()
        return a_.elapsed_time(b_) / steps

    if not args.quick and args.precision == "fp16" and cfg != "cfg5":
        # ---- parity mode: the fp32-grade tensor-core path on the same workload ----
        try:
            n32 = make_net("fp32")
            xs32 = [synth.make_images(B, S, S, seed=100 + i).cuda() for i in range(2)]
            n32.detect_raw(xs32[0], CONF_THR, NMS_THR, False, True, cap)
            n32.freeze_weights()
            _lib.check(lib.yb_set_profiling(n32._ctx, 3), n32._ctx)
            t32 = simple_bench(n32, xs32, lambda i: n32.detect_raw(xs32[i & 1], CONF_THR, NMS_THR, False, True, cap), args.steps)
            lib.yb_get_section_ms(n32._ctx, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
            ach = flops / (a.value * 1e-3) / 1e12
            line["parity_mode"] = {
                "precision": "fp32 (YB_MODE_FP32_TC: fp16 hi/lo operand pairs, 3 tcgen05 MMAs per k-step, two-level fp32 accumulation)",
                "value": B / (t32 * 1e-3), "unit": "images/sec", "ms_per_step": t32, "conv_ms_per_step": a.value,
                "roofline": {"bound": "tensor", "achieved_algorithmic_TFLOPs": ach, "executed_mma_TFLOPs": 3 * ach * (1 - 0.0036),
                             "peak": pk["burst"], "frac_algorithmic": ach / pk["burst"], "frac_executed": 3 * ach * (1 - 0.0036) / pk["burst"],
                             "note": "three fp16 MMAs per algorithmic MAC (the stem, 0.36 % of the FLOPs, runs exact fp32 FMAs on the CUDA cores)"}}
            del n32, xs32
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            line["parity_mode"] = {"error": f"{type(e).__name__}: {e}"}
            torch.cuda.synchronize()

        # ---- BASELINE configs 1, 2, 4: short runs so that the driver's line carries them too ----
        other = {}
        try:    # cfg1: 416x416 single image, forward + decode + NMS -- latency (launch-bound: 80 launches for 66 GFLOP)
            n1 = make_net("fp16", 416)
            x1 = synth.make_images(1, 416, 416, seed=1).cuda()
            n1.detect_raw(x1, CONF_THR, NMS_THR, False, True, cap)
            n1.freeze_weights()
            t1 = simple_bench(n1, x1, lambda i: n1.detect_raw(x1, CONF_THR, NMS_THR, False, True, cap), 50)
            f1 = topology.conv_flops(416, 416)
            other["cfg1_416_b1_detect"] = {"ms_per_image": t1, "value": 1e3 / t1, "unit": "images/sec", "conv_TFLOPs": f1 / (t1 * 1e-3) / 1e12,
                                           "note": "single image: latency of 80 dependent launches, not a throughput configuration"}
            del n1, x1
        except Exception as e:  # noqa: BLE001
            other["cfg1_416_b1_detect"] = {"error": f"{type(e).__name__}: {e}"}
            torch.cuda.synchronize()
        try:    # cfg2: Darknet-53 backbone only, 256x256, batch 64 (tensor roofline)
            n2 = make_net("fp16", 256)
            x2 = [synth.make_images(64, 256, 256, seed=20 + i).cuda() for i in range(2)]
            n2.backbone(x2[0])
            n2.freeze_weights()
            t2 = simple_bench(n2, x2, lambda i: n2.backbone(x2[i & 1]), 20)
            f2 = topology.conv_flops(256, 256, backbone_only=True) * 64
            other["cfg2_backbone_256_b64"] = {"ms_per_step": t2, "value": 64e3 / t2, "unit": "images/sec", "achieved_TFLOPs": f2 / (t2 * 1e-3) / 1e12,
                                              "peak": pk["burst"], "frac": f2 / (t2 * 1e-3) / 1e12 / pk["burst"], "bound": "tensor",
                                              "note": "includes the NHWC->NCHW fp32 conversion of the [64,1024,8,8] output"}
            del n2, x2
        except Exception as e:  # noqa: BLE001
            other["cfg2_backbone_256_b64"] = {"error": f"{type(e).__name__}: {e}"}
            torch.cuda.synchronize()
        try:    # cfg4: NMS stress, 608x608 batch 64, conf 0.001 (~10 k candidates per image) on synthetic head maps
            from yolo_v3_b200.yololayer import decode_heads
            logits = [l.cuda() for l in synth.make_head_logits(64, 608, 608, 80, seed=7)]
            det4 = decode_heads(logits, (608, 608))
            n4 = det4.shape[1]
            cap4 = n4
            res4 = postprocessing_raw(det4, 80, 0.001, NMS_THR, False, True, cap4)
            cand4 = float(res4[3].float().mean())
            surv4 = float(res4[1].float().mean())
            t4 = simple_bench(None, None, lambda i: postprocessing_raw(det4, 80, 0.001, NMS_THR, False, True, cap4), 10)
            by4 = 64 * n4 * 85 * 4
            other["cfg4_nms_stress_608_b64_conf0.001"] = {
                "ms_per_step": t4, "value": 64e3 / t4, "unit": "images/sec", "candidates_per_image": cand4, "survivors_per_image": surv4,
                "first_pass_bytes": by4, "achieved_GBps_whole_call": by4 / (t4 * 1e-3) / 1e9, "peak_GBps": pk["hbm"],
                "frac_whole_call": by4 / (t4 * 1e-3) / 1e9 / pk["hbm"], "bound": "hbm (first pass) then sort / IOU",
                "note": "postprocessing() alone on a resident [64,22743,85] tensor; the whole call (score, scan, sort, NMS, emit) is charged to the first pass's bytes"}
            del logits, det4, res4
        except Exception as e:  # noqa: BLE001
            other["cfg4_nms_stress_608_b64_conf0.001"] = {"error": f"{type(e).__name__}: {e}"}
            torch.cuda.synchronize()
        line["other_configs"] = other
        torch.cuda.empty_cache()

    # ---- CPU baseline: the reference on the host cores, same protocol as --impl reference ----
    cpu_steps = 3 if args.quick else args.cpu_steps
    line["cpu_baseline"], _ = time_cpu(S, REF_IMAGES_PER_STEP, cpu_steps, 2)
    emit(line)
    if args.layers:
        for i, v in enumerate(layer_ms):
            print(f"# op {i:2d} {v:8.4f} ms", file=sys.stderr)
    if world > 1:
        dist.destroy_process_group()


_JSON_FD = None


def emit(line):
    """The one JSON line goes to the process's original stdout; everything else that lands on fd 1 (NCCL prints a
    version banner there when NCCL_DEBUG is set) has been routed to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg3", choices=["cfg3", "cfg5"],
                    help="cfg3: 608x608 batch 32 per GPU (the metric, weak scaling); cfg5: global batch 256 sharded over the ranks (strong scaling)")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--size", type=int, default=608)
    ap.add_argument("--precision", default="fp16", choices=["fp16", "fp32", "fp32_simt"])
    ap.add_argument("--sustained", type=int, default=300, help="steps of the sustained leg (0 = skip)")
    ap.add_argument("--parity-images", type=int, default=8)
    ap.add_argument("--cpu-steps", type=int, default=20, help="steps of 4 images of the cpu_baseline leg")
    ap.add_argument("--quick", action="store_true", help="skip parity, parity_mode and the other configs; 3 CPU steps")
    ap.add_argument("--layers", action="store_true", help="print per-launch device times of the conv section to stderr")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
