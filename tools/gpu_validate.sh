#!/bin/bash
# What the driver runs at round end, in one gpurun call:  gpurun --timeout 1800 -- 'bash tools/gpu_validate.sh <tag>'
#   pytest -m gpu, smoke(), bench.py (default flags) and the reference arm; outputs under gpurun_out/<tag>_*.
T=${1:-val}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -s > gpurun_out/${T}_pytest_full.log 2>&1
grep -E "passed|failed|^FAILED|^ERROR|fp16 deviation|fp16 vs oracle|split mode" gpurun_out/${T}_pytest_full.log | tee gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${T}_smoke.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${T}_bench.err
python - <<PY
import json
d = json.load(open('gpurun_out/${T}_bench.json'))
print('value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1), 'launches', d['gpu_launches'],
      'frac', round(d['roofline']['frac'], 4), 'sustained', round(d['roofline'].get('sustained', {}).get('value', 0), 1),
      'parity_mode', round((d.get('parity_mode') or {}).get('value', 0), 1))
for k in ('e2e_u8_frames', 'e2e_f16_input'):
    print(k, round(d.get(k, {}).get('value', 0), 1))
print('other', {k: (round(v.get('value', 0), 1) if isinstance(v, dict) else v) for k, v in (d.get('other_configs') or {}).items()})
print('parity fp16 set_iou', (d.get('parity') or {}).get('fp16', {}).get('set_iou'), 'fp32_tc logits', (d.get('parity') or {}).get('fp32_tc', {}).get('logits_max_abs_diff_over_max_logit'))
print('cpu', d.get('cpu_baseline', {}).get('value'), d.get('cpu_baseline', {}).get('kind'))
PY
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${T}_bench_reference.json; tail -c 400 gpurun_out/${T}_bench_reference.json
