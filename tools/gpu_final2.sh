#!/bin/bash
# Final validation of the session: what the driver runs (pytest -m gpu, smoke, bench both arms) + memcheck of the new paths.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/final2_pytest.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/final2_smoke.log
echo "== bench (default flags)"; timeout 600 python bench.py > gpurun_out/final2_bench.json 2> gpurun_out/final2_bench.err; tail -c 400 gpurun_out/final2_bench.json; wc -l gpurun_out/final2_bench.json
echo "== bench --impl reference"; timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/final2_bench_ref.json 2>/dev/null; tail -c 300 gpurun_out/final2_bench_ref.json; wc -l gpurun_out/final2_bench_ref.json
echo "== compute-sanitizer memcheck: pair-mode / split-epilogue / halo layer cases + stem"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_fp16.py -x -q -k "(test_tc_layer_vs_torch and (4-4-304 or 6-9-104 or 6-1-9-38 or 6-3-5-76 or 2-4-152 or 74-8-76 or 3-2-10-76 or 1-2-20-152)) or (test_tc_stem_vs_torch and 3-17-23)" 2>&1 | tail -6 | tee gpurun_out/final2_memcheck.log
