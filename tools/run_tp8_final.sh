#!/bin/bash
# Lean multi-GPU stage (gpurun --gpus 8 -- bash tools/run_tp8_final.sh; N=4 gpurun --gpus 4 -- ... for four): all-reduce parity at 4 and 8 ranks,
# the collectives in isolation, the TP=8 bench line and SURVEY 8d config 4.  Every step under its own timeout.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-8}
T=tests/test_gpu_allreduce.py
if [ "$N" = 8 ]; then SEL="$T::test_nvlink_allreduce_matches_nccl_and_host_sum[8-] $T::test_nvlink_allreduce_matches_nccl_and_host_sum[4-] $T::test_nvlink_allgather_lastdim_is_a_bit_exact_cat[8]"
else SEL="$T::test_nvlink_allreduce_matches_nccl_and_host_sum[4-] $T::test_nvlink_allreduce_matches_nccl_and_host_sum[2-twoshot]"; fi
timeout 300 python -m pytest $SEL \
    -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_allreduce_n8.log 2>&1
echo "pytest allreduce rc=$? : $(tail -1 gpurun_out/pytest_allreduce_n8.log)"
grep -E "Error|assert|FAILED" gpurun_out/pytest_allreduce_n8.log | head -10 | cut -c1-300
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $((29330 + RANDOM % 200)) tools/ar_bench.py > gpurun_out/ar_bench_n$N.log 2>&1
echo "ar_bench rc=$?"; grep -E "all-reduce|fused|consumer|all-gather|greedy" gpurun_out/ar_bench_n$N.log
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $((29530 + RANDOM % 200)) bench.py --gpus $N --steps 20 --warmup 3 --skip-cpu-baseline --no-ttft \
    > gpurun_out/bench_tp$N.json 2> gpurun_out/bench_tp$N.err
echo "bench tp$N rc=$? $(tail -1 gpurun_out/bench_tp$N.json | head -c 260)"; tail -2 gpurun_out/bench_tp$N.err | cut -c1-300
[ "$N" = 8 ] && timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $((29130 + RANDOM % 200)) bench.py --gpus $N --model llama3-70b --batch 32 --seqlen 4096 --quant gptq \
    --steps 10 --warmup 3 --skip-cpu-baseline --no-ttft > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err
echo "bench cfg4 rc=$? $(tail -1 gpurun_out/bench_cfg4.json | head -c 300)"; tail -3 gpurun_out/bench_cfg4.err | cut -c1-300
