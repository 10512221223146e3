#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"stem_rows" -s 3 -c 1 -f -o /tmp/r1l_full python tools/one_step.py --steps 1 --warmup 3 --recipe calibrated > /dev/null 2>&1
ncu -i /tmp/r1l_full.ncu-rep --page raw --csv > gpurun_out/r1l_full_raw_stem_rows.csv 2>/dev/null
ncu -i /tmp/r1l_full.ncu-rep --page source --csv > gpurun_out/r1l_src_stem_rows.csv 2>/dev/null
ls -la gpurun_out/r1l*
