#!/bin/bash
# GPU visit "r1m": final validation of the round's last commit + ncu evidence of the halo-tile kernel.
bash tools/gpu_final.sh
echo "== ncu: launch list of one step and full sets of the halo kernel (layers 1, 3, 6)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 246 -c 82 --csv --log-file gpurun_out/r1m_launches.csv \
    python tools/one_step.py --steps 1 --warmup 3 --recipe calibrated > gpurun_out/r1m_launches.log 2>&1
for L in 1 3 6; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv_tc|stem_tc|conv_halo" -s $((225 + L)) -c 1 -f -o /tmp/r1m_full_$L \
      python tools/one_step.py --steps 1 --warmup 3 --recipe calibrated > /dev/null 2>&1
  ncu -i /tmp/r1m_full_$L.ncu-rep --page raw --csv > gpurun_out/r1m_full_raw_layer$L.csv 2>/dev/null
done
ls -la gpurun_out | grep r1m
