"""Times single convolution layers at their real BASELINE shapes (608x608, batch 32) through
yb_run_layer, optionally sweeping the kernel's tuning overrides (YB_TC_BN / YB_TC_STAGES / YB_TC_RING /
YB_TC_BRES), one subprocess per configuration.  Usage on the GPU box:
    python tools/layer_bench.py --layers 1,3,6,11,45 --sweep "YB_TC_BN=128,256;YB_TC_STAGES=3,6"
"""
import argparse
import ctypes
import itertools
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def input_hw(specs, size):
    out, cur = {}, size
    for i, e in enumerate(specs):
        k = e["key"]
        if k.startswith("feature."):
            out[i] = cur
            if e["stride"] == 2:
                cur //= 2
        elif k.startswith("pre_det1") or k.startswith("up1"):
            out[i] = size // 32
        elif k.startswith("pre_det2") or k.startswith("up2"):
            out[i] = size // 16
        else:
            out[i] = size // 8
    return out


def run(layers, batch, size, reps, b2b=1):
    import torch
    from yolo_v3_b200 import _lib, synth, topology
    specs = topology.layer_specs(80)
    hw = input_hw(specs, size)
    sd = synth.make_state_dict(seed=1234, recipe="analytic")
    lib = _lib.load()
    ctx = _lib.create_ctx(0, 80, None)
    for k, v in sd.items():
        if "num_batches" in k:
            continue
        v = v.contiguous()
        _lib.check(lib.yb_set_tensor(ctx, k.encode(), ctypes.c_void_p(v.data_ptr()), v.numel(), 1), ctx)
    _lib.check(lib.yb_finalize(ctx, _lib.YB_MODE_FP16), ctx)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for li in layers:
        e = specs[li]
        h = hw[li]
        ho = h // e["stride"]
        x = torch.rand(batch, 3, h, h, device="cuda") if li == 0 else torch.randn(batch, h, h, e["cin"], device="cuda").half()
        head = not e["bn"]
        cs = (e["cout"] + 15) // 16 * 16 if head else e["cout"]
        out = torch.empty(batch, ho, ho, cs, device="cuda", dtype=torch.float32 if head else torch.float16)
        res = torch.randn(batch, ho, ho, e["cout"], device="cuda").half() if e["res2"] else None
        rp = ctypes.c_void_p(res.data_ptr()) if res is not None else None
        for _ in range(3):
            _lib.check(lib.yb_run_layer(ctx, li, ctypes.c_void_p(x.data_ptr()), batch, h, h, rp, ctypes.c_void_p(out.data_ptr()), st), ctx)
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush.zero_()                              # evict L2 between repetitions
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _k in range(b2b):
                _lib.check(lib.yb_run_layer(ctx, li, ctypes.c_void_p(x.data_ptr()), batch, h, h, rp, ctypes.c_void_p(out.data_ptr()), st), ctx)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) / b2b)
        ts.sort()
        ms = ts[len(ts) // 2]
        fl = 2.0 * batch * ho * ho * e["cout"] * e["cin"] * e["ks"] ** 2
        byts = batch * (h * h * e["cin"] * 2 + ho * ho * cs * (4 if head else 2) * (2 if res is not None else 1))
        print(f"layer {li:2d} {e['key']:26s} {e['cin']:4d}->{e['cout']:4d} k{e['ks']} s{e['stride']} @{h:3d}: {ms:7.4f} ms  "
              f"{fl / ms / 1e9:7.1f} TFLOP/s  {byts / ms / 1e6:7.1f} GB/s(alg)", flush=True)
    lib.yb_destroy(ctx)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", default="1,2,3,4,6,10,11,26,28,43,44,45,61,69,74")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--size", type=int, default=608)
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--sweep", default="")
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--b2b", type=int, default=1, help="launches per timed region (back to back)")
    a = ap.parse_args()
    layers = [int(v) for v in a.layers.split(",")]
    if a.child or not a.sweep:
        run(layers, a.batch, a.size, a.reps, a.b2b)
        sys.exit(0)
    axes = []
    for part in a.sweep.split(";"):
        k, vs = part.split("=")
        axes.append([(k, v) for v in vs.split(",")])
    for combo in itertools.product(*axes):
        env = dict(os.environ)
        env.update({k: v for k, v in combo})
        print("== " + " ".join(f"{k}={v}" for k, v in combo), flush=True)
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child", "--layers", a.layers, "--batch", str(a.batch),
                        "--size", str(a.size), "--reps", str(a.reps), "--b2b", str(a.b2b)], env=env, timeout=600)
