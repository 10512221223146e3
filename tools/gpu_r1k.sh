#!/bin/bash
# GPU visit "r1k": pixel-row stem -- tests, timing against the im2col-row stem.
mkdir -p gpurun_out
echo "### stem tests"; timeout 600 python -m pytest tests/test_gpu_fp16.py -m gpu -q -x -k "test_tc_stem_vs_torch" 2>&1 | tail -12 | tee gpurun_out/r1k_pytest.log
echo "### layer bench: pixel-row stem"; timeout 300 python tools/layer_bench.py --layers 0 2>&1 | tee gpurun_out/r1k_stem_rows.txt
echo "### layer bench: im2col-row stem"; YB_STEM_ROWS=0 timeout 300 python tools/layer_bench.py --layers 0 2>&1 | tee gpurun_out/r1k_stem_tc.txt
