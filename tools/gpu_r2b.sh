#!/bin/bash
# Round 2, second GPU visit: first run of the fp32-grade tensor-core mode (YB_MODE_FP32_TC).
mkdir -p gpurun_out
T=r02b
echo "== split-mode tests"
timeout 900 python -m pytest tests/test_gpu_split.py -q -s -m gpu 2>&1 | grep -v "^$" | tail -60 | tee gpurun_out/${T}_split_pytest.log
echo "== fp32 network tests of the parity suite (now on the split mode)"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "fp32 or dropin or detect_fused" 2>&1 | tail -15 | tee gpurun_out/${T}_parity_fp32_pytest.log
echo "== bench, precision fp32 (split mode)"
timeout 600 python bench.py --precision fp32 --steps 20 --warmup 3 --layers > gpurun_out/${T}_bench_fp32.json 2> gpurun_out/${T}_bench_fp32.err
tail -85 gpurun_out/${T}_bench_fp32.err
python - <<PY
import json; d=json.load(open('gpurun_out/${T}_bench_fp32.json'))
print('fp32 split: value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'conv ms', round(d['roofline']['conv_ms_per_step'],3), 'dets', d['detections_last_step'])
PY
