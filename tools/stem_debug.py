"""Debug aid: run the stem (layer 0) once through yb_run_layer, compare with torch, print the watchdog words on failure."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.nn.functional as F
from yolo_v3_b200 import _lib, synth, topology
B, H, W = [int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (2, 40, 56))]
sd = synth.make_state_dict(seed=1234, recipe="analytic")
lib = _lib.load(); ctx = _lib.create_ctx(0, 80, None)
for k, v in sd.items():
    if "num_batches" in k: continue
    v = v.contiguous()
    _lib.check(lib.yb_set_tensor(ctx, k.encode(), ctypes.c_void_p(v.data_ptr()), v.numel(), 1), ctx)
_lib.check(lib.yb_finalize(ctx, _lib.YB_MODE_FP16), ctx)
rs = np.random.RandomState(7)
x = torch.from_numpy(rs.rand(B, 3, H, W).astype(np.float32))
xd = x.cuda(); out = torch.full((B, H, W, 32), float("nan"), device="cuda", dtype=torch.float16)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
rc = lib.yb_run_layer(ctx, 0, ctypes.c_void_p(xd.data_ptr()), B, H, W, None, ctypes.c_void_p(out.data_ptr()), st)
print("rc", rc)
try:
    torch.cuda.synchronize()
except Exception as e:
    print("sync failed:", str(e).splitlines()[0])
    w = (ctypes.c_int * 8)(); lib.yb_debug_words(ctx, w, 8); print("watchdog words", list(w)); sys.exit(1)
k = "feature.mlist.0"
wt = sd[k + ".conv.weight"].half().float()
y = F.conv2d(x.half().float(), wt, None, 1, 1)
inv = 1.0 / torch.sqrt(sd[k + ".bn.running_var"] + 1e-5); al = inv * sd[k + ".bn.weight"]; be = sd[k + ".bn.bias"] - sd[k + ".bn.running_mean"] * al
ref = F.leaky_relu(y * al.view(1, -1, 1, 1) + be.view(1, -1, 1, 1), 0.1).permute(0, 2, 3, 1)
got = out.float().cpu()
err = (got - ref).abs()
print("nan count", int(torch.isnan(got).sum()), "max err", float(err[~torch.isnan(err)].max()) if (~torch.isnan(err)).any() else None, "ref max", float(ref.abs().max()))
bad = (err > 3e-3 * ref.abs().max() + 2e-3 * ref.abs()) | torch.isnan(got)
print("bad", int(bad.sum()), "of", bad.numel())
if bad.any():
    idx = bad.nonzero()[:10]
    for i in idx: print(tuple(int(v) for v in i), float(got[tuple(i)]), float(ref[tuple(i)]))
