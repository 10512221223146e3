#!/bin/bash
mkdir -p gpurun_out
{
echo "### default"; timeout 300 python tools/layer_bench.py --layers 2,5,10,27,44
echo "### ring 4"; YB_TC_RING=4 timeout 300 python tools/layer_bench.py --layers 2,5,10,27,44
echo "### ring 4, srel 2"; YB_TC_RING=4 YB_TC_SREL=2 timeout 300 python tools/layer_bench.py --layers 2,5,10,27,44
echo "### kps 1 stages 4 (L2)"; YB_TC_STAGES=4 timeout 300 python tools/layer_bench.py --layers 2,5
} 2>&1 | tee gpurun_out/r1r_ring.txt
