#!/bin/bash
# GPU visit "r1h": halo-tile kernel -- layer tests, then per-layer timing against the im2col kernel.
mkdir -p gpurun_out
echo "### halo layer tests"; timeout 600 python -m pytest tests/test_gpu_fp16.py -m gpu -q -x -k "test_tc_layer_vs_torch" 2>&1 | tail -12 | tee gpurun_out/r1h_pytest.log
if grep -q "passed" gpurun_out/r1h_pytest.log && ! grep -q "failed\|error" gpurun_out/r1h_pytest.log; then
  echo "### layer bench: halo"; timeout 300 python tools/layer_bench.py --layers 3,6,8 2>&1 | tee gpurun_out/r1h_layers_halo.txt
  echo "### layer bench: im2col"; YB_HALO=0 timeout 300 python tools/layer_bench.py --layers 3,6,8 2>&1 | tee gpurun_out/r1h_layers_im2col.txt
  echo "### fp16 tests (all)"; timeout 600 python -m pytest tests/test_gpu_fp16.py -m gpu -q -x 2>&1 | tail -4
  echo "### bench"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r1h_bench.json 2> gpurun_out/r1h_bench.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/r1h_bench.json"))
print(round(d["value"],1), "img/s", round(d["ms_per_step"],3), "ms; e2e", round(d["e2e"]["value"],1), "conv", round(d["roofline"]["conv_ms_per_step"],3), "frac", round(d["roofline"]["frac"],3))
PY
else
  python - <<'PY'
import ctypes, sys
sys.path.insert(0, ".")
PY
fi
