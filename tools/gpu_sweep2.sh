#!/bin/bash
{
echo "### cta2 on/off"; timeout 600 python tools/layer_bench.py --layers 9,10,11,27,28,44,45,58,60,66,74 --sweep "YB_TC_CTA2=0,1"
echo "### kps"; timeout 600 python tools/layer_bench.py --layers 1,3,5,10,27,44,68 --sweep "YB_TC_KPS=1,2,4"
echo "### bres"; timeout 300 python tools/layer_bench.py --layers 1,2,3,5,10,68,70 --sweep "YB_TC_BRES=0,1"
echo "### BN for 1x1 and heads"; timeout 600 python tools/layer_bench.py --layers 27,44,52,58,60,66,74 --sweep "YB_TC_BN=64,128,256"
} 2>&1 | tee gpurun_out/sweep2.log
