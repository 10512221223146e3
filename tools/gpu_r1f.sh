#!/bin/bash
# GPU visit "r1f": letterbox v3 + evaluate.py writer tests, bench, times of the other BASELINE configurations.
mkdir -p gpurun_out
echo "### pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r1f_pytest.log
echo "### smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "### bench"; timeout 600 python bench.py > gpurun_out/r1f_bench.json 2> gpurun_out/r1f_bench.err; tail -c 2600 gpurun_out/r1f_bench.json; tail -2 gpurun_out/r1f_bench.err
echo "### other configs"; timeout 900 python tools/config_times.py 2>&1 | tee gpurun_out/r1f_configs.txt
echo "### letterbox ncu"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread
timeout 300 ncu --metrics $M --clock-control none -k regex:"letterbox" -s 3 -c 1 --csv --log-file gpurun_out/r1f_letterbox_metrics.csv python tools/one_letterbox.py > /dev/null 2>&1
grep -v "^==" gpurun_out/r1f_letterbox_metrics.csv | cut -d, -f13- | tail -8
