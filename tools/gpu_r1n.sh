#!/bin/bash
mkdir -p gpurun_out
echo "### fp16 + parity subset"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
echo "### bench"; timeout 600 python bench.py > gpurun_out/r1n_bench.json 2> gpurun_out/r1n_bench.err; wc -l gpurun_out/r1n_bench.json; tail -3 gpurun_out/r1n_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r1n_bench.json"))
print(round(d["value"],1), "img/s", round(d["ms_per_step"],3), "ms; e2e", round(d["e2e"]["value"],1), "conv", round(d["roofline"]["conv_ms_per_step"],3), "frac", round(d["roofline"]["frac"],3))
print({k:(round(v["ms"],4) if isinstance(v,dict) else v) for k,v in d["roofline_hbm"].items()})
PY
