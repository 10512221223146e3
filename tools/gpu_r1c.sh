#!/bin/bash
# GPU visit "r1c": full GPU test suite (letterbox, notebook post-process, fused detect, new stem), stem timing,
# bench with the fused detect path vs the two-kernel path, launch list of one step.
mkdir -p gpurun_out
echo "### pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/r1c_pytest.log
echo "### smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r1c_smoke.log
echo "### stem"; timeout 300 python tools/layer_bench.py --layers 0,1,2 2>&1 | tee gpurun_out/r1c_stem.log
echo "### bench (fused detect, decode v2)"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r1c_bench_new.json 2> gpurun_out/r1c_bench_new.err; tail -c 2600 gpurun_out/r1c_bench_new.json; tail -3 gpurun_out/r1c_bench_new.err
echo "### bench (two-kernel detect, decode v1)"; YB_FUSED_DETECT=0 YB_DECODE_V2=0 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r1c_bench_old.json 2> gpurun_out/r1c_bench_old.err; tail -c 1500 gpurun_out/r1c_bench_old.json
echo "### launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 243 -c 81 --csv --log-file gpurun_out/r1c_launches.csv \
    python tools/one_step.py --steps 1 --warmup 3 --recipe calibrated > gpurun_out/r1c_launches.log 2>&1
tail -3 gpurun_out/r1c_launches.log
