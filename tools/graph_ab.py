import ctypes, sys, time
sys.path.insert(0, '.')
import torch
from yolo_v3_b200 import YoloNet, synth, _lib
sd = synth.make_state_dict(seed=1234, recipe="calibrated")
lib = _lib.load()
for prec in ("fp16", "fp32"):
  for B, S in ((1, 416), (4, 416), (1, 608), (8, 608), (32, 608)):
    if prec == "fp32" and B > 4: continue
    net = YoloNet((S, S), precision=prec); net.load_state_dict(sd); net = net.cuda().eval()
    x = synth.make_images(B, S, S, seed=1).cuda()
    for mode in ("never", "always"):
        net.set_graph_mode(mode)
        net.freeze_weights()
        for _ in range(5): net.detect_raw(x, 0.5, 0.4, False, True, 512)
        torch.cuda.synchronize()
        r0 = lib.yb_graph_replays(net._ctx)
        n = 200 if B <= 8 else 50
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); a.record()
        for _ in range(n): net.detect_raw(x, 0.5, 0.4, False, True, 512)
        b.record(); torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / n * 1e3
        print(f"{prec} B={B} {S}x{S} graph={mode:6s}: {a.elapsed_time(b) / n:.4f} ms/call (device events), {wall:.4f} ms wall, replays {lib.yb_graph_replays(net._ctx) - r0}", flush=True)
