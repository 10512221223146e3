#!/bin/bash
# ncu evidence for one detect step (608x608, batch 32, fp16).  The .ncu-rep files stay on the box (too big
# for gpurun_out); CSV exports come back:
#   <tag>_launches.csv      every launch of one step with its device time (cold cache, serialised)
#   <tag>_conv_metrics.csv  per-launch DRAM bytes / tensor-pipe / L2 metrics of the 75 convolution launches
#   <tag>_full_raw.csv      ncu --set full of five representative convolution launches (raw page)
TAG=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 246 -c 82 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tools/one_step.py --steps 1 --warmup 3 > gpurun_out/${TAG}_launches.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__registers_per_thread
ncu --metrics $M --clock-control none -k regex:"conv_tc|stem_tc" -s 225 -c 75 --csv --log-file gpurun_out/${TAG}_conv_metrics.csv \
    python tools/one_step.py --steps 1 --warmup 3 > gpurun_out/${TAG}_conv_metrics.log 2>&1
ncu --metrics $M --clock-control none -k regex:"decode|pp_" -s 21 -c 7 --csv --log-file gpurun_out/${TAG}_post_metrics.csv \
    python tools/one_step.py --steps 1 --warmup 3 > gpurun_out/${TAG}_post_metrics.log 2>&1
# full set on 5 representative conv launches of the 4th step: layer 1, 6, 11, 28, 45 (launch index = layer index)
for L in 1 6 11 28 45; do
  ncu --set full --clock-control none --import-source on -k regex:"conv_tc|stem_tc" -s $((225 + L)) -c 1 -f -o /tmp/${TAG}_full_$L \
      python tools/one_step.py --steps 1 --warmup 3 > /dev/null 2>&1
  ncu -i /tmp/${TAG}_full_$L.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw_layer$L.csv 2>/dev/null
done
ncu -i /tmp/${TAG}_full_45.ncu-rep --page source --csv > gpurun_out/${TAG}_full_src_layer45.csv 2>/dev/null
du -sh gpurun_out
