#!/bin/bash
# Round 2, first GPU visit: validate the opt-in instantiations written at the end of round 1 (uniform role dispatch,
# leaner stem), settle the host-side experiments, and measure how the tcgen05 accumulator rounds (input to the split mode).
mkdir -p gpurun_out
T=r02a
echo "== accumulator probe"; timeout 300 python tools/tc_accum_probe.py 2>&1 | tee gpurun_out/${T}_tc_accum_probe.txt
echo "== the whole GPU suite with YB_TC_UW=1 YB_STEM_V2=1"
YB_TC_UW=1 YB_STEM_V2=1 timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/${T}_uw_pytest.log
{ echo "### default"; timeout 200 python tools/layer_bench.py --layers 0,1,2,3,6,10,27,44,74
  echo "### YB_TC_UW=1 YB_STEM_V2=1"; YB_TC_UW=1 YB_STEM_V2=1 timeout 200 python tools/layer_bench.py --layers 0,1,2,3,6,10,27,44,74; } 2>&1 | tee gpurun_out/${T}_uw_layers.txt
for i in 1 2; do for uw in 0 1; do
  YB_TC_UW=$uw YB_STEM_V2=$uw timeout 300 python bench.py --steps 100 --warmup 10 > /tmp/uw.json 2>/dev/null
  python -c "
import json; d=json.load(open('/tmp/uw.json')); print('UW/V2=$uw run $i:', round(d['value'],1), 'img/s', round(d['ms_per_step'],4), 'ms/step  conv', round(d['roofline']['conv_ms_per_step'],4), 'dets', d['detections_last_step'], 'e2e', round(d['e2e']['value'],1), 'e2e_u8', round(d['e2e_u8_frames'].get('value',0),1))"
done; done 2>&1 | tee gpurun_out/${T}_uw_ab.txt
echo "== pinned parameter staging"
YB_PINNED_PARAMS=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "letterbox or resize or correct_yolo_boxes or eval_json" 2>&1 | tail -3 | tee gpurun_out/${T}_pinned_pytest.log
YB_PINNED_PARAMS=1 YB_INPUT_F16=1 timeout 600 python bench.py > /tmp/pin.json 2>/dev/null; python -c "
import json; d=json.load(open('/tmp/pin.json')); print('YB_PINNED_PARAMS=1: e2e_u8', round(d['e2e_u8_frames']['value'],1), 'e2e', round(d['e2e']['value'],1), 'e2e_f16', round(d.get('e2e_f16_input',{}).get('value',0),1), 'value', round(d['value'],1))" | tee gpurun_out/${T}_pinned_bench.txt
echo "== L2 persistence of layer outputs, A/B at 100 steps"
for mb in 0 48 80; do
  YB_L2_PERSIST=$mb timeout 300 python bench.py --steps 100 --warmup 10 > /tmp/l2.json 2>/dev/null
  python -c "
import json; d=json.load(open('/tmp/l2.json')); print('YB_L2_PERSIST=$mb:', round(d['value'],1), 'img/s', round(d['ms_per_step'],4), 'ms/step  conv', round(d['roofline']['conv_ms_per_step'],4), 'dets', d['detections_last_step'])"
done 2>&1 | tee gpurun_out/${T}_l2_persist_ab.txt
echo "== blocked-layout access pattern on the 1x1 layers (timing experiment, results wrong)"
{ echo "### NHWC (default)"; timeout 200 python tools/layer_bench.py --layers 5,10,27,44,68
  echo "### YB_TC_EXP_BLOCKED=1"; YB_TC_EXP_BLOCKED=1 timeout 200 python tools/layer_bench.py --layers 5,10,27,44,68; } 2>&1 | tee gpurun_out/${T}_blocked_layout.txt
du -sh gpurun_out
