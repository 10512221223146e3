#!/bin/bash
# ncu / sanitizer evidence for the CURRENT kernel set (608x608, batch 32):  gpurun --timeout 2400 -- 'bash tools/gpu_profile.sh <tag>'
#   <tag>_launches.csv       every launch of one detect step with its device time (cold cache, serialised)
#   <tag>_conv_metrics.csv   per-launch DRAM bytes / tensor-pipe / L2 metrics of the 75 convolution launches (bench.py reads
#                            profiles/r02_conv_metrics.csv for roofline.traffic)
#   <tag>_post_metrics.csv   the decode / post-process kernels
#   <tag>_split_launches.csv the same launch list for precision=fp32 (YB_MODE_FP32_TC)
#   <tag>_full_raw_layer*.csv ncu --set full (raw page) of representative launches
#   <tag>_sanitizer.txt      compute-sanitizer memcheck / racecheck / synccheck on small shapes of both tensor-core modes
T=${1:-prof}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 249 -c 83 --csv --log-file gpurun_out/${T}_launches.csv \
    python tools/one_step.py --steps 1 --warmup 3 --recipe calibrated > gpurun_out/${T}_launches.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__registers_per_thread
ncu --metrics $M --clock-control none -k regex:"conv_tc|stem_block|stem_halo|conv_halo" -s 222 -c 74 --csv --log-file gpurun_out/${T}_conv_metrics.csv \
    python tools/one_step.py --steps 1 --warmup 3 --recipe calibrated > gpurun_out/${T}_conv_metrics.log 2>&1
ncu --metrics $M --clock-control none -k regex:"probe_cells|score_list|decode|pp_" -s 21 -c 7 --csv --log-file gpurun_out/${T}_post_metrics.csv \
    python tools/one_step.py --steps 1 --warmup 3 --recipe calibrated > gpurun_out/${T}_post_metrics.log 2>&1
ncu --metrics gpu__time_duration.sum,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 252 -c 84 --csv \
    --log-file gpurun_out/${T}_split_launches.csv python tools/one_step.py --steps 1 --warmup 3 --recipe calibrated --precision fp32 > gpurun_out/${T}_split_launches.log 2>&1
# conv launch k of a step: 0 = the fused stem + layer 1, k >= 1 = layer k + 1.  Captured: fused first kernel, 64->128 (halo pairs),
# 128->256 3x3 @76, a 1x1 at 38^2, 512->1024 @19, head at 76^2 -- file names keep the LAYER index
for L in 0 6 11 27 45 74; do
  K=$((L == 0 ? 0 : L - 1))
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv_tc|stem_block|stem_halo|conv_halo" -s $((222 + K)) -c 1 -f -o /tmp/${T}_full_$L \
      python tools/one_step.py --steps 1 --warmup 3 --recipe calibrated > /dev/null 2>&1
  ncu -i /tmp/${T}_full_$L.ncu-rep --page raw --csv > gpurun_out/${T}_full_raw_layer$L.csv 2>/dev/null
done
# one big split-mode layer (512->1024 3x3 at 19^2 is conv_tc launch 45 of a split-mode step: the stem is a CUDA-core kernel)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv_tc" -s $((3 * 74 + 44)) -c 1 -f -o /tmp/${T}_full_split45 \
    python tools/one_step.py --steps 1 --warmup 3 --recipe calibrated --precision fp32 > /dev/null 2>&1
ncu -i /tmp/${T}_full_split45.ncu-rep --page raw --csv > gpurun_out/${T}_full_raw_split_layer45.csv 2>/dev/null
{
for tool in memcheck racecheck synccheck; do
  for prec in fp16 fp32; do
    echo "### compute-sanitizer --tool $tool, precision $prec, 2 x 96 x 160"
    timeout 420 compute-sanitizer --tool $tool --print-limit 5 python tools/one_step.py --steps 1 --warmup 0 --batch 2 --size 160 --recipe analytic --precision $prec 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|done|Error|hazard" | head -8
  done
done
# the fused first kernel is not on the 160 x 160 path (its last column tile would be < 80 % full): cover it separately
for tool in memcheck racecheck synccheck; do
  echo "### compute-sanitizer --tool $tool: fused stem + layer 1 kernel, 2 x 40 x 64"
  timeout 250 compute-sanitizer --tool $tool --print-limit 5 python tools/stem_block_debug.py 2 40 64 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|bad |Error|hazard" | head -6
done
} | tee gpurun_out/${T}_sanitizer.txt
du -sh gpurun_out
