#!/bin/bash
# One GPU-box visit: parity tests, tensor-core probe, then (only if the probe is clean) fp16 tests + bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== parity (fp32 / decode / postprocess)" 
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -40 | tee gpurun_out/parity.log
echo "== tc probe"
timeout 1200 python tools/tc_probe.py 2>&1 | tee gpurun_out/tc_probe.log
if grep -q "tc_probe: 15/15" gpurun_out/tc_probe.log; then
  echo "== fp16 tests"
  timeout 900 python -m pytest tests/test_gpu_fp16.py -m gpu -q -s 2>&1 | tail -40 | tee gpurun_out/fp16.log
  echo "== bench"
  timeout 900 python bench.py --steps 10 --warmup 3 --layers > gpurun_out/bench.json 2> gpurun_out/bench.err
  tail -c 3000 gpurun_out/bench.json; tail -n 90 gpurun_out/bench.err
fi
