#!/bin/bash
# GPU visit "r1g" (2 GPUs): batch-sharded bench through torchrun (NCCL weight broadcast + detection all-gather), the
# reference arm under torchrun, and the UMMA row-shift hardware probe for the next round's early-layer design.
mkdir -p gpurun_out
echo "### 2-GPU bench"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 \
    > gpurun_out/r1g_bench_2gpu.json 2> gpurun_out/r1g_bench_2gpu.err
tail -c 1800 gpurun_out/r1g_bench_2gpu.json; tail -3 gpurun_out/r1g_bench_2gpu.err
echo "### reference arm under torchrun"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 3 \
    > gpurun_out/r1g_bench_ref_2gpu.json 2> /dev/null
tail -c 600 gpurun_out/r1g_bench_ref_2gpu.json
echo "### UMMA row-shift probe"
timeout 120 tools/probes/umma_shift_probe.bin 2>&1 | tee gpurun_out/r1g_umma_shift_probe.txt
