#!/bin/bash
# A/B runs that need the tuning overrides: builds the EXPERIMENTS flavour of the library on the box, runs, nothing is kept.
mkdir -p gpurun_out
T=${1:-ab}
make -C yolo_v3_b200/csrc clean > /dev/null; make -C yolo_v3_b200/csrc -j16 EXPERIMENTS=1 2>&1 | grep -E "error|Error" | head
python - <<'PY' 2>&1 | tee gpurun_out/${T}_graph_cfg1.txt
import ctypes, sys, time
sys.path.insert(0, '.')
import torch
from yolo_v3_b200 import YoloNet, synth, _lib
sd = synth.make_state_dict(seed=1234, recipe="calibrated")
lib = _lib.load()
for B, S in ((1, 416), (4, 416), (1, 608), (8, 608)):
    net = YoloNet((S, S), precision="fp16"); net.load_state_dict(sd); net = net.cuda().eval()
    x = synth.make_images(B, S, S, seed=1).cuda()
    for mode in ("never", "always"):
        net.set_graph_mode(mode)
        net.freeze_weights()
        for _ in range(5): net.detect_raw(x, 0.5, 0.4, False, True, 512)
        torch.cuda.synchronize()
        r0 = lib.yb_graph_replays(net._ctx)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); a.record()
        for _ in range(200): net.detect_raw(x, 0.5, 0.4, False, True, 512)
        b.record(); torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / 200 * 1e3
        print(f"B={B} {S}x{S} graph={mode:6s}: {a.elapsed_time(b) / 200:.4f} ms/call (device events), {wall:.4f} ms wall, replays {lib.yb_graph_replays(net._ctx) - r0}", flush=True)
PY
for i in 1 2; do for v in 0 1; do
  YB_UP_DIRECT=$v timeout 300 python bench.py --quick --steps 100 --warmup 10 --sustained 0 > /tmp/ab.json 2>/dev/null
  python -c "
import json; d=json.load(open('/tmp/ab.json')); print('YB_UP_DIRECT=$v run $i:', round(d['value'],1), 'img/s', round(d['ms_per_step'],4), 'ms/step  conv', round(d['roofline']['conv_ms_per_step'],4), 'dets', d['detections_last_step'])"
done; done 2>&1 | tee gpurun_out/${T}_up_direct_ab.txt
