#!/bin/bash
# Steady-state A/B of the epilogue back-off (YB_TC_EPI_SLEEP ns) with power / clock sampling: 300-step runs.
mkdir -p gpurun_out
for rep in 1 2; do
for S in 0 200 1000; do
  nvidia-smi --query-gpu=power.draw,clocks.sm --format=csv,noheader,nounits -lms 100 > gpurun_out/power2_s${S}_$rep.csv 2>/dev/null &
  SMI=$!
  YB_TC_EPI_SLEEP=$S timeout 600 python bench.py --steps 300 --warmup 20 > gpurun_out/power2_bench_s${S}_$rep.json 2>/dev/null
  kill $SMI 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/power2_bench_s${S}_$rep.json"))
rows=[l.strip().split(", ") for l in open("gpurun_out/power2_s${S}_$rep.csv") if l.strip()]
busy=[(float(r[0]), float(r[1])) for r in rows if float(r[0])>600]
pw=sorted(p for p,_ in busy); ck=sorted(c for _,c in busy)
print("sleep=$S rep=$rep:", round(d["value"],1), "img/s", round(d["ms_per_step"],3), "ms | median W", pw[len(pw)//2] if pw else None, "median MHz", ck[len(ck)//2] if ck else None)
PY
done
done
