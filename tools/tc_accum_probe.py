"""How does the tcgen05 fp32 accumulator round?  (design input for the fp32-grade split mode)

Runs the 1x1 head convolution pre_det1.mlist.6 (1024 -> 255, fp32 output = the raw TMEM accumulator + 0 bias) through
yb_run_layer in YB_MODE_FP16 on operands that are exactly representable in fp16, and compares with the exact fp64 result:
the products are exact in fp32, so the whole difference is the accumulation (64 MMA steps of K = 16 per output).
Prints max / rms relative error and the mean SIGNED error in units of fp32 ulp for (a) mixed-sign and (b) all-positive
operands -- a systematic negative mean in (b) means the accumulator truncates instead of rounding to nearest.
"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from yolo_v3_b200 import _lib, synth, topology

specs = topology.layer_specs(80)
li = next(i for i, e in enumerate(specs) if e["key"] == "pre_det1.mlist.6")
lib = _lib.load()
for name, positive in (("mixed sign", False), ("all positive", True)):
    sd = synth.make_state_dict(seed=1234, recipe="analytic")
    rs = np.random.RandomState(5)
    w = rs.standard_normal((255, 1024, 1, 1)).astype(np.float32) * 0.05
    if positive:
        w = np.abs(w)
    w = torch.from_numpy(w).half().float()
    sd["pre_det1.mlist.6.weight"] = w
    sd["pre_det1.mlist.6.bias"] = torch.zeros(255)
    ctx = _lib.create_ctx(0, 80, None)
    for k, v in sd.items():
        if "num_batches" in k:
            continue
        v = v.contiguous()
        _lib.check(lib.yb_set_tensor(ctx, k.encode(), ctypes.c_void_p(v.data_ptr()), v.numel(), 1), ctx)
    _lib.check(lib.yb_finalize(ctx, _lib.YB_MODE_FP16), ctx)
    B, h = 4, 19
    x = torch.from_numpy(rs.standard_normal((B, h, h, 1024)).astype(np.float32))
    if positive:
        x = x.abs()
    x = x.half()
    out = torch.empty(B, h, h, 256, device="cuda", dtype=torch.float32)
    xd = x.cuda()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.yb_run_layer(ctx, li, ctypes.c_void_p(xd.data_ptr()), B, h, h, None, ctypes.c_void_p(out.data_ptr()), st), ctx)
    torch.cuda.synchronize()
    got = out.cpu()[..., :255].double().reshape(-1, 255)
    exact = x.double().reshape(-1, 1024) @ w.double().reshape(255, 1024).t()
    f32 = (x.float().reshape(-1, 1024) @ w.float().reshape(255, 1024).t()).double()      # CPU fp32 GEMM, for scale
    ulp = np.spacing(np.abs(exact.numpy()).astype(np.float32)).astype(np.float64)
    e_tc = (got - exact).numpy() / ulp
    e_cpu = (f32 - exact).numpy() / ulp
    print(f"{name:13s} K=1024: tensor core  max|err|={np.abs(e_tc).max():8.2f} ulp  rms={np.sqrt((e_tc**2).mean()):7.3f} ulp  mean signed={e_tc.mean():+8.3f} ulp"
          f"   |  CPU fp32 GEMM  max={np.abs(e_cpu).max():8.2f}  rms={np.sqrt((e_cpu**2).mean()):7.3f}  mean signed={e_cpu.mean():+8.3f}")
    s = np.sqrt((exact.numpy() ** 2).mean())
    print(f"{'':13s} relative to rms|y|={s:.3f}: tensor core max {np.abs((got - exact).numpy()).max() / s:.3e}, CPU fp32 max {np.abs((f32 - exact).numpy()).max() / s:.3e}")
    lib.yb_destroy(ctx)
