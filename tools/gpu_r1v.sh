#!/bin/bash
# r1v: halo kernel in CTA-pair mode for the 64->128 layers at 152x152
mkdir -p gpurun_out
echo "== layer tests"; timeout 600 python -m pytest tests/test_gpu_fp16.py -x -q -k "test_tc_layer_vs_torch" 2>&1 | tail -6 | tee gpurun_out/r1v_pytest.log
{ echo "### pair mode"; timeout 200 python tools/layer_bench.py --layers 6,8
  echo "### YB_HALO_PAIR=0"; YB_HALO_PAIR=0 timeout 200 python tools/layer_bench.py --layers 6,8; } 2>&1 | tee gpurun_out/r1v_layers.txt
echo "== bench"; timeout 600 python bench.py > gpurun_out/r1v_bench.json 2> gpurun_out/r1v_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r1v_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['detections_last_step'])"
