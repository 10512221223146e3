#!/usr/bin/env python
"""bench.py -- images/sec of the YOLOv3 detect hot path at 608x608, batch 32 per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = one pass of the hot path over one batch of synthetic images: 75 fused convolutions
(tcgen05 tensor cores, fp16 in / fp32 accumulate) -> 3-scale anchor decode -> box filter / sort /
IOU / greedy NMS (yb_detect), i.e. what test.py:35-36 of the reference computes per batch
(net(imgs) + postprocessing(cat(det), conf 0.5, nms 0.4)).  Weights are random-init (BN-calibrated,
seeded) of the real architecture; images are uniform noise -- there is no network for datasets.

Printed JSON (one line, rank 0):
  value     whole-job images/sec with the input batches already resident in HBM (CUDA events on the
            launch stream, max over ranks);
  e2e       the same through the public Python API with HOST buffers: every step copies its batch
            from pinned host memory and reads the detections back (copy engine overlapped with
            compute by double buffering; both inside the timed region).  The fp32 batch is 142 MB per
            step, so the PCIe link (~26 GB/s on the test boxes) caps this figure near 6 000 img/s per GPU;
  e2e_u8_frames  the same from uint8 camera frames ([B,480,640,3], 29.5 MB per step): H2D copy, letterbox
            on the device (yb_letterbox), detect, detections back -- all inside the timed region;
  roofline  the convolution stack (conv_tc_kernel / conv_halo_kernel, 74 launches per step + the stem):
            algorithmic 2*MAC FLOPs / CUDA-event time of the conv section, against the measured
            sustained bf16 tensor peak of MEASURED_PEAKS.json;
  cpu_baseline  the CPU oracle port (torch fp32 oneDNN convs + the reference's NMS algorithm) on a
            bounded sample of the same workload, all host threads.

--impl reference times that CPU path alone (the reference is pure Python/PyTorch and cannot travel
to the GPU box; oracle/ is its faithful restatement).
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


def ncu_conv_traffic():
    """DRAM bytes (read+write) of the 75 convolution launches of one step, from the committed ncu capture
    profiles/r01final_conv_metrics.csv (608x608 batch 32 only)."""
    import csv
    p = os.path.join(ROOT, "profiles", "r01final_conv_metrics.csv")
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
    """Sustained dense bf16 TFLOP/s and HBM copy GB/s out of the driver-written MEASURED_PEAKS.json, whose exact key
    names this repository does not control: known names first, then any numeric leaf whose path says what it is (a
    sustained tensor figure is preferred over a burst one -- the conv stack is timed inside a long step).  Returns
    (tensor, hbm), either may be None."""
    flat = list(_flatten(d))
    by = dict(flat)
    tensor = by.get("bf16_tflops_sustained")
    hbm = by.get("hbm_gbs")
    if tensor is None:
        cand = [(k, v) for k, v in flat if any(t in k for t in ("bf16", "tflop", "tf/s", "tensor")) and 100.0 <= v <= 5000.0]
        sus = [v for k, v in cand if "sustain" in k]
        other = [v for k, v in cand if "burst" not in k and "peak" not in k]
        tensor = sus[0] if sus else (other[0] if other else (min(v for _, v in cand) if cand else None))
    if hbm is None:
        cand = [v for k, v in flat if any(t in k for t in ("hbm", "dram", "copy", "gb/s", "gbs", "gbps", "bandwidth")) and 500.0 <= v <= 20000.0]
        hbm = cand[0] if cand else None
    return tensor, hbm


def peaks():
    """Roofline denominators: MEASURED_PEAKS.json when the driver has written it, else the profiling recipe's fallback."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    tensor = hbm = None
    if os.path.exists(p):
        try:
            tensor, hbm = parse_peaks(json.load(open(p)))
        except Exception as e:  # noqa: BLE001 -- a malformed file must not take the bench down
            print(f"[bench] MEASURED_PEAKS.json unreadable ({e}); using the fallback peaks", file=sys.stderr)
    if tensor is None and hbm is None:
        return dict(tensor=1400.0, hbm=6650.0, src="fallback")
    src = "measured" if tensor is not None and hbm is not None else "measured+fallback"
    return dict(tensor=tensor if tensor is not None else 1400.0, hbm=hbm if hbm is not None else 6650.0, src=src)


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([v.strip() for v in out.split(",")])
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(self.samples)}


def cpu_reference(batch, hw, reps, num_threads=None):
    """The reference's CPU path (oracle port): forward + decode + postprocessing, images/sec."""
    import torch
    from oracle import yolo_oracle as O
    from yolo_v3_b200 import synth
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core
    torch.set_num_threads(num_threads or os.cpu_count() or 1)
    sd = synth.make_state_dict(seed=1234, recipe="calibrated")
    x = synth.make_images(batch, hw, hw, seed=0)
    best = None
    O.postprocessing(torch.cat(O.forward(sd, x[:1]), 1), 80, CONF_THR, NMS_THR)       # warm-up
    for _ in range(reps):
        t0 = time.perf_counter()
        det = torch.cat(O.forward(sd, x), 1)
        t1 = time.perf_counter()
        O.postprocessing(det, 80, CONF_THR, NMS_THR)
        t2 = time.perf_counter()
        if best is None or t2 - t0 < best[0]:
            best = (t2 - t0, t1 - t0, t2 - t1)
    return dict(value=batch / best[0], unit="images/sec", cores=torch.get_num_threads(), kind="port",
                sample=f"{batch} images {hw}x{hw}, best of {reps}: forward+decode {best[1]:.2f}s, postprocessing {best[2]:.3f}s "
                       f"(torch {torch.__version__} fp32, oneDNN; oracle/yolo_oracle.py restates darknet.py/yololayer.py/utils.py)")


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; the reference itself is a directory of Python scripts
    that does not travel to the GPU box) timed for W warm-up + exactly K steps, each step a bounded sample of
    --ref-batch images of the same workload, all host threads; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import yolo_oracle as O
    from yolo_v3_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    sd = synth.make_state_dict(seed=1234, recipe="calibrated")
    x = synth.make_images(args.ref_batch, args.size, args.size, seed=0)

    def step():
        det = torch.cat(O.forward(sd, x), 1)
        return O.postprocessing(det, 80, CONF_THR, NMS_THR)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = args.ref_batch * args.steps / dt
    base = dict(value=value, unit="images/sec", cores=torch.get_num_threads(), kind="port",
                sample=f"{args.steps} steps x {args.ref_batch} images {args.size}x{args.size} in {dt:.1f}s after {args.warmup} warm-up steps "
                       f"(torch {torch.__version__} fp32 oneDNN convs + the reference's NMS algorithm; oracle/yolo_oracle.py "
                       "restates darknet.py/yololayer.py/utils.py)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "images/sec", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"yolov3_{args.size}x{args.size}_b{args.batch}_detect", "conf_thr": CONF_THR, "nms_thr": NMS_THR,
                       "note": f"CPU path; each step is a bounded sample of {args.ref_batch} images of the batch-{args.batch} workload"},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def run_ours(args):
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
    B, S = args.batch, args.size
    N = topology.num_boxes(S, S)

    sd = synth.make_state_dict(seed=1234, recipe="calibrated")
    net = YoloNet((S, S), precision=args.precision)
    net.load_state_dict(sd)
    net = net.cuda().eval()
    lib = _lib.load()

    # two resident input batches (each 142 MB > the 126 MB L2), alternated between steps
    xs = [synth.make_images(B, S, S, seed=100 * rank + i).cuda() for i in range(2)]
    cap = 512
    stream = torch.cuda.current_stream()

    comm = None
    if world > 1:
        from yolo_v3_b200 import parallel
        comm = parallel.DetectionGather(net, rank, world, B, cap)
        net(xs[0][:1])                      # creates + finalises the engine
        comm.broadcast_weights()

    def step(i):
        rows, counts, src, cand = net.detect_raw(xs[i & 1], CONF_THR, NMS_THR, False, True, cap)
        if comm is not None:
            return comm.allgather(rows, counts)
        return rows, counts

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    ctx = net._ctx
    launches0 = lib.yb_launch_count(ctx)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        out = step(i)
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.yb_launch_count(ctx) - launches0
    counts_h = out[1].cpu()
    assert int(counts_h.max()) <= cap, "detection capacity overflow in the timed region"
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())

    # ---- end to end: pinned host batches in, detections out, every step ----
    hx = [synth.make_images(B, S, S, seed=7 + i).pin_memory() for i in range(2)]
    dx = [torch.empty_like(xs[0]) for _ in range(2)]
    copy_stream = torch.cuda.Stream()
    ready = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]
    h_rows = torch.empty(B * (world if comm else 1), cap, 7).pin_memory()
    h_counts = torch.empty(B * (world if comm else 1), dtype=torch.int32).pin_memory()

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
        e2e_loop(max(2, args.warmup), host, devb, to_input)
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

    e2e_ms = time_e2e(hx, dx, lambda x: x)
    del hx, dx

    # ---- optional legs (one GPU only, and never allowed to take the contract line down) ----
    # (a) from camera frames: pinned uint8 [B,480,640,3] in, letterbox on the device (yb_letterbox, the N1 row), detect,
    #     detections out.  The fp32 batch above is 142 MB per step, i.e. the PCIe link (~26 GB/s measured) caps it near
    #     6 000 img/s per GPU whatever the kernels do; frames are 29.5 MB per step.
    # (b) from fp16 images read by the stem directly (yb_set_input_dtype): bit-identical detections, half the PCIe bytes.
    FH, FW = 480, 640
    extra = {}
    if world == 1:
        d2h = int(h_rows.numel() * 4 + h_counts.numel() * 4)
        try:
            import numpy as np
            from yolo_v3_b200.utils import letterbox_batch
            frames = np.stack([synth.make_photo(FH, FW, 90 + i) for i in range(B)])
            hu = [torch.from_numpy(frames).pin_memory(), torch.from_numpy(np.ascontiguousarray(frames[::-1])).pin_memory()]
            du = [torch.empty(B, FH, FW, 3, dtype=torch.uint8, device=dev) for _ in range(2)]
            t_ms = time_e2e(hu, du, lambda u: letterbox_batch(list(u), (S, S))[0])
            del hu, du
            extra["e2e_u8_frames"] = {
                "value": B * args.steps / (t_ms * 1e-3), "unit": "images/sec", "h2d_bytes_per_step": B * FH * FW * 3,
                "d2h_bytes_per_step": d2h, "ms_per_step": t_ms / args.steps,
                "input": f"uint8 [B,{FH},{FW},3] frames, letterboxed to {S}x{S} on the device (yb_letterbox) inside the timed region"}
        except Exception as e:  # noqa: BLE001
            extra["e2e_u8_frames"] = {"error": f"{type(e).__name__}: {e}"}
            torch.cuda.synchronize()
        # (b) is timed on request (YB_INPUT_F16=1) until this leg itself has run on a GPU box at the full batch size
        if args.precision == "fp16" and os.environ.get("YB_INPUT_F16") == "1":
            try:
                hh = [synth.make_images(B, S, S, seed=7 + i).half().pin_memory() for i in range(2)]
                dh = [torch.empty(B, 3, S, S, dtype=torch.float16, device=dev) for _ in range(2)]
                t_ms = time_e2e(hh, dh, lambda x: x)
                del hh, dh
                extra["e2e_f16_input"] = {
                    "value": B * args.steps / (t_ms * 1e-3), "unit": "images/sec", "h2d_bytes_per_step": B * 3 * S * S * 2,
                    "d2h_bytes_per_step": d2h, "ms_per_step": t_ms / args.steps,
                    "input": f"fp16 [B,3,{S},{S}] read by the stem directly (yb_set_input_dtype; bit-identical detections)"}
            except Exception as e:  # noqa: BLE001
                extra["e2e_f16_input"] = {"error": f"{type(e).__name__}: {e}"}
                torch.cuda.synchronize()

    # ---- roofline: section times of the same step, CUDA events per section on the launch stream ----
    import ctypes
    # events at the section boundaries, recorded without synchronising so that the steps still run back to back as in
    # the timed loop above; one query at the end averages them
    _lib.check(lib.yb_set_profiling(ctx, 3), ctx)
    reps = min(args.steps, 16)
    for i in range(3):
        net.detect_raw(xs[i & 1], CONF_THR, NMS_THR, False, True, cap)
    a, b, c = ctypes.c_float(), ctypes.c_float(), ctypes.c_float()
    lib.yb_get_section_ms(ctx, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))      # discard the warm-up sets
    for i in range(reps):
        net.detect_raw(xs[i & 1], CONF_THR, NMS_THR, False, True, cap)
    lib.yb_get_section_ms(ctx, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
    conv_ms, dec_ms, post_ms = a.value, b.value, c.value
    layer_ms = []
    if args.layers:
        _lib.check(lib.yb_set_profiling(ctx, 2), ctx)      # + one event per convolution
        for i in range(3):
            net.detect_raw(xs[i & 1], CONF_THR, NMS_THR, False, True, cap)
            buf = (ctypes.c_float * 80)()
            n = lib.yb_get_layer_ms(ctx, buf, 80)
            cur = [buf[j] for j in range(min(n, 80))]
            layer_ms = cur if not layer_ms else [x + y for x, y in zip(layer_ms, cur)]
        layer_ms = [v / 3 for v in layer_ms]
    _lib.check(lib.yb_set_profiling(ctx, 0), ctx)

    # ---- HBM-bound kernels timed alone (CUDA events on the launch stream, inputs > L2 or alternated) ----
    hbm = {}
    if rank == 0:
        from yolo_v3_b200.utils import letterbox_batch, postprocessing_raw

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

        _lib.check(lib.yb_set_profiling(ctx, 1), ctx)
        d_alone = 0.0
        for i in range(5):                                  # yb_forward: conv stack + standalone decode (writes det)
            dets = net(xs[i & 1], None)
            a, b, c = ctypes.c_float(), ctypes.c_float(), ctypes.c_float()
            lib.yb_get_section_ms(ctx, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
            d_alone += b.value / 5
        _lib.check(lib.yb_set_profiling(ctx, 0), ctx)
        det_cat = net._last_det                             # the concatenated [B,N,85] tensor det1..3 are views of
        p_alone = timed(lambda: postprocessing_raw(det_cat, 80, CONF_THR, NMS_THR, False, True, cap))
        photos = [torch.from_numpy(synth.make_photo(480, 640, 50 + i)).cuda() for i in range(B)]
        l_alone = timed(lambda: letterbox_batch(photos, (S, S)))
        lb_bytes = B * (480 * 640 * 3 + S * S * 12)
        hbm = {"decode_alone_ms": d_alone, "post_alone_ms": p_alone, "letterbox_ms": l_alone, "letterbox_bytes": lb_bytes}
        del det_cat, dets, photos

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    sampler.stop_flag = True
    sampler.join(timeout=2)
    pk = peaks()
    flops = topology.conv_flops(S, S) * B
    achieved = flops / (conv_ms * 1e-3) / 1e12
    row_bytes = B * N * 85 * 4              # one pass over the [B,N,85] fp32 tensor (padded logit pitch ignored)
    total_imgs = B * world * args.steps
    cpu = cpu_reference(args.ref_batch, S, 2)

    def gbps(nbytes, t_ms):
        return {"ms": t_ms, "achieved_GBps": nbytes / (t_ms * 1e-3) / 1e9 if t_ms else None,
                "frac": nbytes / (t_ms * 1e-3) / 1e9 / pk["hbm"] if t_ms else None}
    line = {
        "metric": METRIC, "value": total_imgs / (ms * 1e-3), "unit": "images/sec", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16" if args.precision == "fp16" else "f32", "data": "synthetic",
        "config": {"workload": f"yolov3_{S}x{S}_b{B}_detect", "batch_per_gpu": B, "global_batch": B * world, "img": S,
                   "conf_thr": CONF_THR, "nms_thr": NMS_THR, "precision": args.precision,
                   "l2": "inputs larger than L2 (two alternating 142 MB batches; activations 0.8 GB per layer)",
                   "parallelism": f"batch-sharded x{world}, weights broadcast once, detections all-gathered per step"},
        "e2e": {"value": total_imgs / (e2e_ms * 1e-3), "unit": "images/sec", "h2d_bytes_per_step": B * 3 * S * S * 4,
                "d2h_bytes_per_step": int(h_rows.numel() * 4 + h_counts.numel() * 4), "ms_per_step": e2e_ms / args.steps,
                "input": f"fp32 [B,3,{S},{S}] in [0,1], what the reference's predict() moves with .cuda() (test.py:32)"},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
        "roofline": {"bound": "tensor", "kernel": "conv_tc_kernel + conv_halo_kernel (74 launches/step) + stem_tc_kernel", "achieved": achieved,
                     "peak": pk["tensor"], "unit": "TFLOP/s", "frac": achieved / pk["tensor"],
                     "traffic": ncu_conv_traffic() if (S == 608 and B == 32) else None,
                     "traffic_note": "DRAM read+write bytes of the 75 conv launches of one step (ncu, profiles/r01final_conv_metrics.csv); "
                                     "algorithmic activation+weight bytes: 13.0e9",
                     "peak_source": f"{pk['src']} sustained bf16 (MEASURED_PEAKS.json)",
                     "conv_ms_per_step": conv_ms, "flops_per_step": flops},
        "roofline_hbm": {
            # inside the timed step (yb_detect): decode + score fused, reads the head maps once, det is never written
            "decode_score_fused": dict(gbps(row_bytes, dec_ms), bytes="logits in"),
            "post_scan_sort_nms_emit_ms": post_ms,
            # the same kernels behind the reference's separate calls, timed alone
            "decode": dict(gbps(2 * row_bytes, hbm.get("decode_alone_ms", 0.0)), bytes="logits in + det out"),
            "postprocess": dict(gbps(row_bytes, hbm.get("post_alone_ms", 0.0)), bytes="det in"),
            "letterbox": dict(gbps(hbm.get("letterbox_bytes", 0), hbm.get("letterbox_ms", 0.0)),
                              bytes="uint8 480x640 sources in + fp32 CHW canvases out"),
            "peak_GBps": pk["hbm"]},
        "cpu_baseline": cpu,
        "detections_last_step": int(counts_h.sum()),
    }
    line.update(extra)
    emit(line)
    if args.layers:
        specs = topology.layer_specs(80)
        for i, v in enumerate(layer_ms):
            print(f"# layer {i:2d} {specs[i]['key']:28s} {v:8.4f} ms", file=sys.stderr)
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
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--size", type=int, default=608)
    ap.add_argument("--precision", default="fp16", choices=["fp16", "fp32"])
    ap.add_argument("--ref-batch", type=int, default=0,
                    help="images per CPU-baseline step (bounded sample); default 16 for the cpu_baseline leg, 4 per step for --impl reference")
    ap.add_argument("--layers", action="store_true", help="print per-layer device times to stderr")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.ref_batch <= 0:
        args.ref_batch = 4 if args.impl == "reference" else 16
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
