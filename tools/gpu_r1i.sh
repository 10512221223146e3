#!/bin/bash
# GPU visit "r1i": stride-2 halo kernel (parity planes) -- layer tests, then timing against the im2col kernel.
mkdir -p gpurun_out
echo "### layer tests"; timeout 600 python -m pytest tests/test_gpu_fp16.py -m gpu -q -x -k "test_tc_layer_vs_torch" 2>&1 | tail -12 | tee gpurun_out/r1i_pytest.log
echo "### layer bench: halo s2"; timeout 300 python tools/layer_bench.py --layers 1,3 2>&1 | tee gpurun_out/r1i_layers_halo.txt
echo "### layer bench: im2col s2"; YB_HALO_S2=0 timeout 300 python tools/layer_bench.py --layers 1 2>&1 | tee gpurun_out/r1i_layers_im2col.txt
