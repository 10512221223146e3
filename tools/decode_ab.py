"""Same-box A/B of the stand-alone decode (yb_forward's decode section) -- run with an EXPERIMENTS build:
   YB_DECODE_STAGED=0/1 python tools/decode_ab.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yolo_v3_b200 import YoloNet, synth, _lib
sd = synth.make_state_dict(seed=1234, recipe="calibrated")
net = YoloNet((608, 608)); net.load_state_dict(sd); net = net.cuda().eval()
xs = [synth.make_images(32, 608, 608, seed=i).cuda() for i in range(2)]
lib = _lib.load()
net(xs[0], None)
ctx = net._ctx
_lib.check(lib.yb_set_profiling(ctx, 1), ctx)
a, b, c = ctypes.c_float(), ctypes.c_float(), ctypes.c_float()
ts = []
for i in range(12):
    det = net(xs[i & 1], None)
    lib.yb_get_section_ms(ctx, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
    if i >= 2: ts.append(b.value)
ts.sort()
print(f"YB_DECODE_STAGED={os.environ.get('YB_DECODE_STAGED', '0')}: decode section median {ts[len(ts) // 2]:.4f} ms, min {ts[0]:.4f} ms; checksum {float(torch.cat(det, 1).double().sum()):.6f}")
