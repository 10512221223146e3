#!/bin/bash
# r1w: two epilogue groups (stem: alternate tiles; conv_tc 32-column sub-tiles: alternate tiles / sub-tiles)
mkdir -p gpurun_out
echo "== layer + stem tests"; timeout 600 python -m pytest tests/test_gpu_fp16.py -x -q -k "test_tc_layer_vs_torch or test_tc_stem_vs_torch" 2>&1 | tail -6 | tee gpurun_out/r1w_pytest.log
{ echo "### new"; timeout 200 python tools/layer_bench.py --layers 0,2,58,66,74
  echo "### YB_TC_EPISPLIT=0"; YB_TC_EPISPLIT=0 timeout 200 python tools/layer_bench.py --layers 2,58,66,74; } 2>&1 | tee gpurun_out/r1w_layers.txt
echo "== bench"; timeout 600 python bench.py > gpurun_out/r1w_bench.json 2> gpurun_out/r1w_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r1w_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_u8_frames']['value'], d['roofline']['frac'], d['detections_last_step'])"; tail -3 gpurun_out/r1w_bench.err
