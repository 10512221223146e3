#!/bin/bash
# First GPU visit of the next round: re-validate, refresh the ncu evidence for the kernels that changed at the end of round 1
# (halo pair mode, BN=128 pair mode, split epilogue, stem with two epilogue groups), and settle two open A/Bs.
#   gpurun --timeout 1500 -- 'bash tools/gpu_r2_first.sh'
mkdir -p gpurun_out
T=r02a
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/${T}_pytest.log
echo "== bench (default flags) + the fp16-input e2e leg"
YB_INPUT_F16=1 timeout 600 python bench.py --layers > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<PY
import json; d=json.load(open('gpurun_out/${T}_bench.json'))
print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1),
      'e2e_u8', round(d['e2e_u8_frames']['value'],1), 'e2e_f16', round(d.get('e2e_f16_input',{}).get('value',0),1), 'frac', round(d['roofline']['frac'],4))
PY
echo "== uniform role dispatch (YB_TC_UW) and the leaner stem (YB_STEM_V2): the whole GPU suite with both, then an A/B of the step"
YB_TC_UW=1 YB_STEM_V2=1 timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/${T}_uw_pytest.log
{ echo "### default"; timeout 200 python tools/layer_bench.py --layers 0,1,2,3,6,10,27,74
  echo "### YB_TC_UW=1 YB_STEM_V2=1"; YB_TC_UW=1 YB_STEM_V2=1 timeout 200 python tools/layer_bench.py --layers 0,1,2,3,6,10,27,74; } 2>&1 | tee gpurun_out/${T}_uw_layers.txt
for i in 1 2; do for uw in 0 1; do
  YB_TC_UW=$uw YB_STEM_V2=$uw timeout 300 python bench.py --steps 100 --warmup 10 > /tmp/uw.json 2>/dev/null
  python -c "
import json; d=json.load(open('/tmp/uw.json')); print('UW/V2=$uw run $i:', round(d['value'],1), 'img/s', round(d['ms_per_step'],4), 'ms/step  conv', round(d['roofline']['conv_ms_per_step'],4), 'dets', d['detections_last_step'])"
done; done 2>&1 | tee gpurun_out/${T}_uw_ab.txt
echo "== pinned parameter staging: letterbox / resize / box-correction tests and the uint8-frames e2e with it"
YB_PINNED_PARAMS=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "letterbox or resize or correct_yolo_boxes or eval_json" 2>&1 | tail -3 | tee gpurun_out/${T}_pinned_pytest.log
YB_PINNED_PARAMS=1 timeout 600 python bench.py > /tmp/pin.json 2>/dev/null; python -c "
import json; d=json.load(open('/tmp/pin.json')); print('YB_PINNED_PARAMS=1: e2e_u8', round(d['e2e_u8_frames']['value'],1), 'e2e', round(d['e2e']['value'],1), 'value', round(d['value'],1))" | tee gpurun_out/${T}_pinned_bench.txt
echo "== early weights inside the step, 4 x A/B at 100 steps"
for i in 1 2 3 4; do for be in 0 1; do
  YB_TC_BEARLY=$be timeout 300 python bench.py --steps 100 --warmup 10 > /tmp/ab.json 2>/dev/null
  python -c "
import json; d=json.load(open('/tmp/ab.json')); print('BEARLY=$be run $i:', round(d['value'],1), 'img/s', round(d['ms_per_step'],4), 'ms/step  conv', round(d['roofline']['conv_ms_per_step'],4), 'clocks', d['clocks']['sm_mhz'])"
done; done 2>&1 | tee gpurun_out/${T}_bearly_ab.txt
echo "== L2 persistence of layer outputs (host-side launch attribute), A/B at 100 steps"
for i in 1 2; do for mb in 0 48 80; do
  YB_L2_PERSIST=$mb timeout 300 python bench.py --steps 100 --warmup 10 > /tmp/l2.json 2>/dev/null
  python -c "
import json; d=json.load(open('/tmp/l2.json')); print('YB_L2_PERSIST=$mb run $i:', round(d['value'],1), 'img/s', round(d['ms_per_step'],4), 'ms/step  conv', round(d['roofline']['conv_ms_per_step'],4), 'dets', d['detections_last_step'])"
done; done 2>&1 | tee gpurun_out/${T}_l2_persist_ab.txt
echo "== blocked-layout access pattern on the 1x1 layers (timing experiment, results wrong)"
{ echo "### NHWC (default)"; timeout 200 python tools/layer_bench.py --layers 5,10,27,44,68
  echo "### YB_TC_EXP_BLOCKED=1"; YB_TC_EXP_BLOCKED=1 timeout 200 python tools/layer_bench.py --layers 5,10,27,44,68; } 2>&1 | tee gpurun_out/${T}_blocked_layout.txt
echo "== ncu: launch list, per-launch conv metrics, full sets of the changed kernels"
ncu --metrics gpu__time_duration.sum --clock-control none -s 246 -c 82 --csv --log-file gpurun_out/${T}_launches.csv \
    python tools/one_step.py --steps 1 --warmup 3 --recipe calibrated > gpurun_out/${T}_launches.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__registers_per_thread
ncu --metrics $M --clock-control none -k regex:"conv_tc|stem_tc|conv_halo" -s 225 -c 75 --csv --log-file gpurun_out/${T}_conv_metrics.csv \
    python tools/one_step.py --steps 1 --warmup 3 --recipe calibrated > gpurun_out/${T}_conv_metrics.log 2>&1
# launch index = layer index: stem, 64->32 (split epilogue), 64->128 s2 (BN=128 pairs), 64->128 (halo pairs), a 1x1 at 38^2, up2, head at 76^2
for L in 0 2 4 6 27 67 74; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv_tc|stem_tc|conv_halo" -s $((225 + L)) -c 1 -f -o /tmp/${T}_full_$L \
      python tools/one_step.py --steps 1 --warmup 3 --recipe calibrated > /dev/null 2>&1
  ncu -i /tmp/${T}_full_$L.ncu-rep --page raw --csv > gpurun_out/${T}_full_raw_layer$L.csv 2>/dev/null
done
du -sh gpurun_out
