#!/bin/bash
{
echo "### im2col vs tiled A loads (timing only)"
timeout 300 python tools/layer_bench.py --layers 1,3,4,6,11,28,45 --sweep "YB_TC_EXP_TILED=0,1"
echo "### stem + decode"
timeout 300 python tools/layer_bench.py --layers 0
} 2>&1 | tee gpurun_out/exp.log
