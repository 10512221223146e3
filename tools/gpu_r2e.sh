#!/bin/bash
# Round 2: where does the split mode lose time?  Chunk length, pair mode, and the in-kernel time line.
mkdir -p gpurun_out
T=r02e
run() {  # label, env...
  label=$1; shift
  env "$@" timeout 300 python bench.py --precision fp32 --steps 10 --warmup 3 --layers > /tmp/b.json 2> /tmp/b.err
  python - "$label" <<PY
import json, sys, re
d = json.load(open('/tmp/b.json'))
ls = [float(m.group(1)) for m in re.finditer(r'# layer +\d+ \S+ +([0-9.]+) ms', open('/tmp/b.err').read())]
pick = [1, 3, 6, 11, 28, 45, 10, 27, 44, 58]
print(sys.argv[1], 'img/s', round(d['value'], 1), 'conv ms', round(d['roofline']['conv_ms_per_step'], 3), ' '.join(f'L{i}:{ls[i]:.3f}' for i in pick if i < len(ls)), flush=True)
PY
}
{
run "default          "
run "YB_SPLIT_CHUNK=2 " YB_SPLIT_CHUNK=2
run "YB_SPLIT_CHUNK=4 " YB_SPLIT_CHUNK=4
run "YB_SPLIT_CHUNK=16" YB_SPLIT_CHUNK=16
run "YB_TC_CTA2=0     " YB_TC_CTA2=0
run "YB_TC_BN=64      " YB_TC_BN=64
} 2>&1 | tee gpurun_out/${T}_split_variants.txt
YB_TC_TRACE=1 timeout 300 python tools/one_step.py --precision fp32 --batch 32 --warmup 1 --steps 1 2> gpurun_out/${T}_trace.txt > /dev/null
grep -c "tc trace" gpurun_out/${T}_trace.txt
timeout 600 python -m pytest tests/test_gpu_split.py -q -m gpu 2>&1 | tail -5
