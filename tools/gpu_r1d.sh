#!/bin/bash
# GPU visit "r1d": tests of the round's last kernels (tiled letterbox, resize, IaaLetterbox rule, objectness-first
# scoring), A/B of the scoring and letterbox kernels, bench.
mkdir -p gpurun_out
echo "### pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/r1d_pytest.log
echo "### smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r1d_smoke.log
echo "### bench"; timeout 600 python bench.py > gpurun_out/r1d_bench.json 2> gpurun_out/r1d_bench.err; tail -c 2800 gpurun_out/r1d_bench.json; tail -3 gpurun_out/r1d_bench.err
echo "### bench, cell-at-a-time scoring + direct letterbox"; YB_SCORE_OBJ_FIRST=0 YB_LB_DIRECT=1 timeout 600 python bench.py --steps 10 > gpurun_out/r1d_bench_ab.json 2> gpurun_out/r1d_bench_ab.err; tail -c 1500 gpurun_out/r1d_bench_ab.json
echo "### bench --impl reference"; timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r1d_bench_ref.json 2>/dev/null; tail -c 700 gpurun_out/r1d_bench_ref.json
