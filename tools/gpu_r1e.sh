#!/bin/bash
# GPU visit "r1e": two-kernel scoring (tests + A/B), final ncu evidence of the round's kernel set.
mkdir -p gpurun_out
echo "### pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r1e_pytest.log
for M in 2 1; do
  echo "### bench YB_SCORE_MODE=$M"; YB_SCORE_MODE=$M timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r1e_bench_m$M.json 2> gpurun_out/r1e_bench_m$M.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r1e_bench_m$M.json"))
print(round(d["value"],1), "img/s", round(d["ms_per_step"],3), "ms; e2e", round(d["e2e"]["value"],1), "conv", round(d["roofline"]["conv_ms_per_step"],3), "frac", round(d["roofline"]["frac"],3))
print({k:(round(v["ms"],4) if isinstance(v,dict) else v) for k,v in d["roofline_hbm"].items()})
PY
done
echo "### ncu: launch list of one step (81+ launches)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 246 -c 82 --csv --log-file gpurun_out/r1e_launches.csv \
    python tools/one_step.py --steps 1 --warmup 3 --recipe calibrated > gpurun_out/r1e_launches.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__registers_per_thread
echo "### ncu: decode / scoring / post-process metrics"
timeout 300 ncu --metrics $M --clock-control none -k regex:"probe_cells|score_|decode_|pp_" -s 21 -c 7 --csv --log-file gpurun_out/r1e_post_metrics.csv \
    python tools/one_step.py --steps 1 --warmup 3 --recipe calibrated > /dev/null 2>&1
echo "### ncu: letterbox metrics + full"
timeout 300 ncu --metrics $M --clock-control none -k regex:"letterbox" -s 3 -c 1 --csv --log-file gpurun_out/r1e_letterbox_metrics.csv \
    python tools/one_letterbox.py > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"letterbox" -s 3 -c 1 -f -o /tmp/r1e_full_lb python tools/one_letterbox.py > /dev/null 2>&1
ncu -i /tmp/r1e_full_lb.ncu-rep --page raw --csv > gpurun_out/r1e_full_raw_letterbox.csv 2>/dev/null
echo "### ncu: full of the stem, two convolutions and the scoring kernels"
for L in 0 28 45; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv_tc|stem_tc" -s $((225 + L)) -c 1 -f -o /tmp/r1e_full_$L \
      python tools/one_step.py --steps 1 --warmup 3 --recipe calibrated > /dev/null 2>&1
  ncu -i /tmp/r1e_full_$L.ncu-rep --page raw --csv > gpurun_out/r1e_full_raw_layer$L.csv 2>/dev/null
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"probe_cells|score_list" -s 6 -c 2 -f -o /tmp/r1e_full_score \
    python tools/one_step.py --steps 1 --warmup 3 --recipe calibrated > /dev/null 2>&1
ncu -i /tmp/r1e_full_score.ncu-rep --page raw --csv > gpurun_out/r1e_full_raw_score.csv 2>/dev/null
ls -la gpurun_out | grep r1e
