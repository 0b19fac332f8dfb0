#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port $((29330 + RANDOM % 200)) tools/ar_bench.py > gpurun_out/ar_bench_n2.log 2>&1
echo "ar_bench rc=$?"; grep -E "all-reduce|fused|consumer|all-gather|greedy|phases|entry|stored|complete|row done" gpurun_out/ar_bench_n2.log | cut -c1-160
