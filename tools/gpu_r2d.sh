#!/bin/bash
# Round 2: split mode with two-level accumulation.
mkdir -p gpurun_out
T=r02d
timeout 600 python tools/tc_accum_probe2.py 2>&1 | tee gpurun_out/${T}_split_error_vs_k.txt
echo "== split-mode tests"
timeout 900 python -m pytest tests/test_gpu_split.py -q -s -m gpu > gpurun_out/${T}_split_pytest_full.log 2>&1
grep -E "^layer|split mode|passed|failed|^FAILED|Mismatch|Max abs|Max rel" gpurun_out/${T}_split_pytest_full.log | tee gpurun_out/${T}_split_pytest.log
echo "== fp32 network tests of the parity suite (split mode)"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "fp32 or dropin or detect_fused" 2>&1 | tail -15 | tee gpurun_out/${T}_parity_fp32_pytest.log
echo "== bench, precision fp32 (split mode)"
timeout 600 python bench.py --precision fp32 --steps 20 --warmup 3 --layers > gpurun_out/${T}_bench_fp32.json 2> gpurun_out/${T}_bench_fp32.err
grep "# layer" gpurun_out/${T}_bench_fp32.err | awk '{printf "%s ", $6} END {print ""}'
python - <<PY
import json; d=json.load(open('gpurun_out/${T}_bench_fp32.json'))
print('fp32 split: value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'conv ms', round(d['roofline']['conv_ms_per_step'],3), 'dets', d['detections_last_step'])
PY
