#!/bin/bash
mkdir -p gpurun_out
T=r02c
timeout 600 python tools/tc_accum_probe2.py 2>&1 | tee gpurun_out/${T}_split_error_vs_k.txt
timeout 900 python -m pytest tests/test_gpu_split.py -q -s -m gpu -k "layer or stem" 2>&1 | grep -E "^layer|passed|failed" | tee gpurun_out/${T}_split_layers.log
