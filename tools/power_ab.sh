#!/bin/bash
# Steady-state A/B of the halo-tile kernel with power / clock sampling: 300-step runs (1.7 s of GPU time each),
# nvidia-smi polled in the background.
mkdir -p gpurun_out
for rep in 1 2; do
for H in 1 0; do
  nvidia-smi --query-gpu=power.draw,clocks.sm,clocks_event_reasons.sw_power_cap --format=csv,noheader,nounits -lms 100 > gpurun_out/power_h${H}_$rep.csv 2>/dev/null &
  SMI=$!
  YB_HALO=$H timeout 600 python bench.py --steps 300 --warmup 20 > gpurun_out/power_bench_h${H}_$rep.json 2>/dev/null
  kill $SMI 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/power_bench_h${H}_$rep.json"))
rows=[l.strip().split(", ") for l in open("gpurun_out/power_h${H}_$rep.csv") if l.strip()]
busy=[(float(r[0]), float(r[1])) for r in rows if float(r[0])>600]
pw=sorted(p for p,_ in busy); ck=sorted(c for _,c in busy)
print("halo=$H rep=$rep:", round(d["value"],1), "img/s", round(d["ms_per_step"],3), "ms  conv", round(d["roofline"]["conv_ms_per_step"],3),
      "| samples>600W:", len(busy), "median W", pw[len(pw)//2] if pw else None, "median MHz", ck[len(ck)//2] if ck else None)
PY
done
done
