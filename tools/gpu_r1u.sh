#!/bin/bash
# r1u: division-free tile walkers, BN=128 pair mode with resident half weight slabs (64->128 stride 2), halo epilogue with
# scale/bias in registers.  Validity first (pytest), then per-layer A/B timings, then the driver's own sequence.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r1u_pytest.log
echo "== layer bench (new code)"
{ echo "### new code"; timeout 300 python tools/layer_bench.py --layers 1,2,3,4,5,6,10,11
  echo "### layer 4 without BN=128 pair mode (YB_TC_PAIR128=0)"; YB_TC_PAIR128=0 timeout 300 python tools/layer_bench.py --layers 4
  echo "### layer 4 pair mode, one k-block per stage (YB_TC_KPS=1)"; YB_TC_KPS=1 timeout 300 python tools/layer_bench.py --layers 4
} 2>&1 | tee gpurun_out/r1u_layers.txt
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r1u_smoke.log
echo "== bench (default flags)"; timeout 900 python bench.py > gpurun_out/r1u_bench.json 2> gpurun_out/r1u_bench.err; tail -c 600 gpurun_out/r1u_bench.json; wc -l gpurun_out/r1u_bench.json
echo "== bench --impl reference"; timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r1u_bench_ref.json 2>/dev/null; tail -c 300 gpurun_out/r1u_bench_ref.json; wc -l gpurun_out/r1u_bench_ref.json
echo "== compute-sanitizer memcheck: pair-mode layer cases"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_fp16.py -x -q -k "test_tc_layer_vs_torch and (4-4-304 or 4-7-304 or 6-9-104 or 3-2-10-76 or 1-2-20-152)" 2>&1 | tail -6 | tee gpurun_out/r1u_memcheck.log
