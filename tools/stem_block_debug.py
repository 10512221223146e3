"""Debug aid: run the fused stem + layer-1 kernel once through yb_run_stem_block, compare with torch (same fp16-rounded
operands, fp16-rounded stem output), print the watchdog words on failure; optionally time it at a real shape."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.nn.functional as F
from yolo_v3_b200 import _lib, synth
B, H, W = [int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (2, 40, 64))]
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 0
sd = synth.make_state_dict(seed=1234, recipe="analytic")
lib = _lib.load(); ctx = _lib.create_ctx(0, 80, None)
for k, v in sd.items():
    if "num_batches" in k: continue
    v = v.contiguous()
    _lib.check(lib.yb_set_tensor(ctx, k.encode(), ctypes.c_void_p(v.data_ptr()), v.numel(), 1), ctx)
_lib.check(lib.yb_finalize(ctx, _lib.YB_MODE_FP16), ctx)
rs = np.random.RandomState(7)
x = torch.from_numpy(rs.rand(B, 3, H, W).astype(np.float32))
xd = x.cuda(); out = torch.full((B, H // 2, W // 2, 64), float("nan"), device="cuda", dtype=torch.float16)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
rc = lib.yb_run_stem_block(ctx, ctypes.c_void_p(xd.data_ptr()), B, H, W, ctypes.c_void_p(out.data_ptr()), st)
print("rc", rc, _lib.last_error() if rc and hasattr(_lib, "last_error") else "")
try:
    torch.cuda.synchronize()
except Exception as e:
    print("sync failed:", str(e).splitlines()[0])
    w = (ctypes.c_int * 8)(); lib.yb_debug_words(ctx, w, 8); print("watchdog words [flag, block, role, barrier, parity]", list(w)); sys.exit(1)

def layer(k, xin, stride):
    wt = sd[k + ".conv.weight"].half().float()
    y = F.conv2d(xin, wt, None, stride, 1)
    inv = 1.0 / torch.sqrt(sd[k + ".bn.running_var"] + 1e-5); al = inv * sd[k + ".bn.weight"]; be = sd[k + ".bn.bias"] - sd[k + ".bn.running_mean"] * al
    return F.leaky_relu(y * al.view(1, -1, 1, 1) + be.view(1, -1, 1, 1), 0.1)
if B * H * W <= 4 * 608 * 608:
    y0 = layer("feature.mlist.0", x.half().float(), 1).half().float()
    ref = layer("feature.mlist.1", y0, 2).permute(0, 2, 3, 1)
    got = out.float().cpu()
    err = (got - ref).abs()
    fin = ~torch.isnan(err)
    print("nan count", int(torch.isnan(got).sum()), "max err", float(err[fin].max()) if fin.any() else None, "ref max", float(ref.abs().max()))
    bad = (err > 6e-3 * ref.abs().max() + 4e-3 * ref.abs()) | torch.isnan(got)
    print("bad", int(bad.sum()), "of", bad.numel())
    if bad.any():
        idx = bad.nonzero()
        print("bad rows (y) histogram", torch.bincount(idx[:, 1], minlength=H // 2).tolist()[:40])
        print("bad cols (x) histogram", torch.bincount(idx[:, 2], minlength=W // 2).tolist()[:80])
        print("bad channel histogram", torch.bincount(idx[:, 3], minlength=64).tolist())
        for i in idx[:8]: print(tuple(int(v) for v in i), float(got[tuple(i)]), float(ref[tuple(i)]))
if reps:
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        lib.yb_run_stem_block(ctx, ctypes.c_void_p(xd.data_ptr()), B, H, W, ctypes.c_void_p(out.data_ptr()), st)
        b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); print(f"fused stem + layer 1 @ {B}x{H}x{W}: {ts[len(ts) // 2]:.4f} ms (median of {reps})")
