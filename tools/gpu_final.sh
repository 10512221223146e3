#!/bin/bash
# Final validation on the GPU box: what the driver runs (pytest -m gpu, smoke, bench both arms) + memcheck.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/final_pytest.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/final_smoke.log
echo "== bench (default flags)"; timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 600 gpurun_out/final_bench.json
echo "== bench --impl reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/final_bench_ref.json 2>/dev/null; tail -c 400 gpurun_out/final_bench_ref.json
echo "== compute-sanitizer memcheck (small shapes)"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python - <<'PY' 2>&1 | tail -12 | tee gpurun_out/final_memcheck.log
import torch, sys
sys.path.insert(0, ".")
from yolo_v3_b200 import YoloNet, synth
sd = synth.make_state_dict(seed=1234, recipe="analytic")
for prec, hw in (("fp16", (96, 160)), ("fp32", (64, 64))):
    net = YoloNet(hw, precision=prec); net.load_state_dict(sd); net = net.cuda().eval()
    x = synth.make_images(2, hw[1], hw[0], seed=1).cuda()
    out = net.detect(x, 0.05, 0.4)
    torch.cuda.synchronize()
    print(prec, [tuple(o.shape) for o in out])
print("memcheck run finished")
PY
