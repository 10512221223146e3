#!/bin/bash
# Final validation on the GPU box: what the driver runs (pytest -m gpu, smoke, bench both arms) + memcheck.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/final_pytest.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/final_smoke.log
echo "== bench (default flags)"; timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 700 gpurun_out/final_bench.json; wc -l gpurun_out/final_bench.json
echo "== bench --impl reference"; timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final_bench_ref.json 2>/dev/null; tail -c 400 gpurun_out/final_bench_ref.json; wc -l gpurun_out/final_bench_ref.json
echo "== compute-sanitizer memcheck (small shapes)"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python - <<'PY' 2>&1 | tail -12 | tee gpurun_out/final_memcheck.log
import torch, sys
import numpy as np
sys.path.insert(0, ".")
from yolo_v3_b200 import YoloNet, synth
from yolo_v3_b200.utils import letterbox_batch, resize_batch
from yolo_v3_b200.notebook import postprocessing as nb_post
sd = synth.make_state_dict(seed=1234, recipe="analytic")
for prec, hw in (("fp16", (96, 160)), ("fp16", (64, 608)), ("fp32", (64, 64))):     # 608 wide: the halo-tile kernel runs (304, 152)
    net = YoloNet(hw, precision=prec); net.load_state_dict(sd); net = net.cuda().eval()
    x = synth.make_images(2, hw[1], hw[0], seed=1).cuda()
    out = net.detect(x, 0.05, 0.4)
    ev = net.detect(x, 0.05, 0.45, is_eval=True)
    det = torch.cat(net(x, None), 1)
    nb = nb_post(det, 80, 0.3, 0.4)
    torch.cuda.synchronize()
    print(prec, hw, [tuple(o.shape) for o in out], [tuple(o.shape) for o in ev][:1], [tuple(o.shape) for o in nb][:1])
imgs = [synth.make_photo(97, 131, 1), synth.make_photo(240, 427, 2), synth.make_photo(720, 1280, 3)]
a, t = letterbox_batch(imgs, (160, 160)); b = resize_batch(imgs, (96, 64)); torch.cuda.synchronize()
print("letterbox", tuple(a.shape), "resize", tuple(b.shape))
print("memcheck run finished")
PY
