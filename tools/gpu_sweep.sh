#!/bin/bash
mkdir -p gpurun_out
{
echo "### correctness"; timeout 900 python tools/tc_probe.py | tail -3
echo "### default config"; timeout 600 python tools/layer_bench.py
echo "### stages sweep (latency vs bandwidth bound?)"; timeout 900 python tools/layer_bench.py --layers 11,28,45,61,69 --sweep "YB_TC_STAGES=2,3,4"
echo "### BN sweep"; timeout 900 python tools/layer_bench.py --layers 4,6,11,26,28,43,45,61,69 --sweep "YB_TC_BN=64,128,256"
echo "### b-resident off"; YB_TC_BRES=0 timeout 600 python tools/layer_bench.py --layers 1,2,3,5,10,68,70
echo "### ring sweep"; timeout 600 python tools/layer_bench.py --layers 6,11,28,45 --sweep "YB_TC_RING=2,3,4"
} 2>&1 | tee gpurun_out/sweep.log
