"""Error of the split mode (YB_MODE_FP32_TC) per layer as a function of the accumulation chain length: yb_run_layer on
fp32 inputs against float64, for layers of growing K.  rms / max error relative to rms|ref| and the regression slope of the
error on the value (a truncating accumulator shrinks results toward zero: negative slope growing with the chain length)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F
from yolo_v3_b200 import _lib, synth, topology

specs = topology.layer_specs(80)
lib = _lib.load()
sd = synth.make_state_dict(seed=1234, recipe="calibrated")
ctx = _lib.create_ctx(0, 80, None)
for k, v in sd.items():
    if "num_batches" in k:
        continue
    v = v.contiguous()
    _lib.check(lib.yb_set_tensor(ctx, k.encode(), ctypes.c_void_p(v.data_ptr()), v.numel(), 1), ctx)
_lib.check(lib.yb_finalize(ctx, _lib.YB_MODE_FP32_TC), ctx)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for li in (2, 5, 10, 27, 44, 6, 11, 28, 45):
    e = specs[li]
    for kind in ("normal", "positive"):
        rs = np.random.RandomState(li)
        B, h = 2, 19 * e["stride"]
        x = rs.standard_normal((B, h, h, e["cin"])).astype(np.float32)
        if kind == "positive":
            x = np.abs(x)
        x = torch.from_numpy(x)
        ho = h // e["stride"]
        out = torch.empty(B, ho, ho, e["cout"], device="cuda", dtype=torch.float32)
        xd = x.cuda()
        _lib.check(lib.yb_run_layer(ctx, li, ctypes.c_void_p(xd.data_ptr()), B, h, h, None, ctypes.c_void_p(out.data_ptr()), st), ctx)
        torch.cuda.synchronize()
        k = e["key"]
        raw = F.conv2d(x.double().permute(0, 3, 1, 2), sd[k + ".conv.weight"].double(), None, e["stride"], (e["ks"] - 1) // 2)
        invstd = 1.0 / torch.sqrt(sd[k + ".bn.running_var"] + 1e-5)
        alpha = (invstd * sd[k + ".bn.weight"]).double().view(1, -1, 1, 1)
        beta = (sd[k + ".bn.bias"] - sd[k + ".bn.running_mean"] * (invstd * sd[k + ".bn.weight"])).double().view(1, -1, 1, 1)
        y = out.cpu().double().permute(0, 3, 1, 2)
        # undo the epilogue to look at the raw accumulator: leaky^-1, then (v - beta) / alpha
        v = torch.where(y > 0, y, y / 0.1)
        acc = (v - beta) / alpha
        err = acc - raw
        rms = float(raw.pow(2).mean().sqrt())
        slope = float((err * raw).sum() / (raw * raw).sum())
        f32 = F.conv2d(x.permute(0, 3, 1, 2), sd[k + ".conv.weight"], None, e["stride"], (e["ks"] - 1) // 2).double()
        e32 = f32 - raw
        K = e["cin"] * e["ks"] ** 2
        print(f"layer {li:2d} K={K:5d} ({3 * K // 16:4d} MMAs) {kind:8s}: TC rms {float(err.pow(2).mean().sqrt()) / rms:.2e} max {float(err.abs().max()) / rms:.2e} "
              f"slope {slope:+.2e} | CPU fp32 rms {float(e32.pow(2).mean().sqrt()) / rms:.2e} max {float(e32.abs().max()) / rms:.2e} slope {float((e32 * raw).sum() / (raw * raw).sum()):+.2e}",
              flush=True)
lib.yb_destroy(ctx)
