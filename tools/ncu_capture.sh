#!/bin/bash
# ncu evidence for one detect step (608x608, batch 32, fp16): launch list + full capture of the conv kernel.
# The .ncu-rep stays on the box (too big for gpurun_out); CSV exports come back.
TAG=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 246 -c 82 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tools/one_step.py --steps 1 --warmup 3 > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 222 -c 74 -f -o /tmp/${TAG}_conv \
    python tools/one_step.py --steps 1 --warmup 3 > gpurun_out/${TAG}_conv.log 2>&1
ncu -i /tmp/${TAG}_conv.ncu-rep --page raw --csv > gpurun_out/${TAG}_conv_raw.csv 2>/dev/null
for L in 0 5 10 44; do
  ncu -i /tmp/${TAG}_conv.ncu-rep --page source --csv --launch-skip $L --launch-count 1 > gpurun_out/${TAG}_conv_src_launch${L}.csv 2>/dev/null
done
ls -la /tmp/${TAG}_conv.ncu-rep gpurun_out/
du -sh gpurun_out
