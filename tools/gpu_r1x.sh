#!/bin/bash
# r1x: stem L2 prefetch sweep, direct (upsample) epilogue on sixteen warps
mkdir -p gpurun_out
echo "== fp16 tests"; timeout 600 python -m pytest tests/test_gpu_fp16.py -x -q 2>&1 | tail -4 | tee gpurun_out/r1x_pytest.log
{ for pf in 0 1 2 4 8; do echo "### YB_STEM_PF=$pf"; YB_STEM_PF=$pf timeout 120 python tools/layer_bench.py --layers 0; done
  echo "### upsample layers are timed inside the step (yb_run_layer runs them without the 2x2 replication)"; } 2>&1 | tee gpurun_out/r1x_layers.txt
for pf in 0 2; do echo "== bench YB_STEM_PF=$pf"; YB_STEM_PF=$pf timeout 600 python bench.py --layers > gpurun_out/r1x_bench_pf$pf.json 2> gpurun_out/r1x_bench_pf$pf.err; python -c "
import json; d=json.load(open('gpurun_out/r1x_bench_pf$pf.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_u8_frames']['value'], d['roofline']['frac'], d['detections_last_step'])"; grep -E "layer +(0|59|67) " gpurun_out/r1x_bench_pf$pf.err | head -5; done
