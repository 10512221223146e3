#!/bin/bash
mkdir -p gpurun_out
echo "### 4-GPU bench"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 5 \
    > gpurun_out/r1o_bench_4gpu.json 2> gpurun_out/r1o_bench_4gpu.err
wc -l gpurun_out/r1o_bench_4gpu.json; tail -c 900 gpurun_out/r1o_bench_4gpu.json; tail -2 gpurun_out/r1o_bench_4gpu.err
