#!/bin/bash
# r1z: weights requested before griddepcontrol.wait (YB_TC_BEARLY) inside the step, A/B/A/B on one box
mkdir -p gpurun_out
for i in 1 2; do for be in 0 1; do
  YB_TC_BEARLY=$be timeout 300 python bench.py > gpurun_out/r1z_b$be_$i.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/r1z_b$be_$i.json')); print('BEARLY=$be run $i:', round(d['value'],1), 'img/s', round(d['ms_per_step'],4), 'ms/step  conv', round(d['roofline']['conv_ms_per_step'],4), 'ms  dets', d['detections_last_step'])"
done; done 2>&1 | tee gpurun_out/r1z_bearly_ab.txt
