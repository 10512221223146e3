#!/bin/bash
# GPU visit "r1b": L2 prefetch (A operand / residual / stem input), early weight touch, store-release rule.
# A/B against the previous behaviour through the env overrides, then bench + ncu of the stem and layer 2.
mkdir -p gpurun_out
OLD="YB_TC_PF=0 YB_TC_BEARLY=0 YB_TC_SREL=1 YB_STEM_PF=0"
L="0,1,2,3,4,5,6,9,10,11,27,28,44,45,58,59,60,66,68,74"
{
echo "### fp16 kernel tests (new defaults)"; timeout 600 python -m pytest tests/test_gpu_fp16.py -m gpu -q -x 2>&1 | tail -5
echo "### OLD"; env $OLD timeout 300 python tools/layer_bench.py --layers $L
echo "### NEW"; timeout 300 python tools/layer_bench.py --layers $L
echo "### PF sweep"; timeout 600 python tools/layer_bench.py --layers 1,2,3,4,5,6,10,27,60,68 --sweep "YB_TC_PF=0,1,4,8"
echo "### SREL sweep"; timeout 400 python tools/layer_bench.py --layers 1,2,3,4,6,10,11,27 --sweep "YB_TC_SREL=0,1,2"
echo "### BEARLY off"; YB_TC_BEARLY=0 timeout 300 python tools/layer_bench.py --layers 2,5,10,27,44,45
echo "### stem PF sweep"; timeout 300 python tools/layer_bench.py --layers 0 --sweep "YB_STEM_PF=0,2,8,16"
} 2>&1 | tee gpurun_out/r1b_sweep.log
echo "### bench OLD"; env $OLD timeout 600 python bench.py --steps 10 --warmup 3 --layers > gpurun_out/r1b_bench_old.json 2> gpurun_out/r1b_bench_old.err; tail -c 1500 gpurun_out/r1b_bench_old.json
echo "### bench NEW"; timeout 600 python bench.py --steps 10 --warmup 3 --layers > gpurun_out/r1b_bench_new.json 2> gpurun_out/r1b_bench_new.err; tail -c 1500 gpurun_out/r1b_bench_new.json
for Lx in 0 2; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv_tc|stem_tc" -s $((225 + Lx)) -c 1 -f -o /tmp/r1b_full_$Lx \
      python tools/one_step.py --steps 1 --warmup 3 > /dev/null 2>&1
  ncu -i /tmp/r1b_full_$Lx.ncu-rep --page raw --csv > gpurun_out/r1b_full_raw_layer$Lx.csv 2>/dev/null
done
ls -la gpurun_out | head -30
