#!/bin/bash
# Multi-GPU stage (gpurun --gpus N -- bash tools/run_tp.sh N [quick]): the all-reduce / all-gather /
# sharded-argmax tests at every world size the box offers, then bench.py under torchrun, the in-graph
# timeline of rank 0, and the C++-only decode demo (one process, one worker thread per GPU).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
if [ "${SKIP_TESTS:-0}" != 1 ]; then
  timeout 1500 python -m pytest tests/test_gpu_allreduce.py -m gpu -q --tb=short -p no:cacheprovider \
      > gpurun_out/pytest_allreduce_n$N.log 2>&1
  echo "pytest allreduce rc=$? : $(tail -3 gpurun_out/pytest_allreduce_n$N.log | tr '\n' ' ')"
  grep -E "Error|assert|FAILED" gpurun_out/pytest_allreduce_n$N.log | head -10
fi
run() {  # tag, env assignments...
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port $((29530 + RANDOM % 200)) bench.py --gpus $N --steps 20 --warmup 3 --skip-cpu-baseline \
      > gpurun_out/bench_tp${N}_$tag.json 2> gpurun_out/bench_tp${N}_$tag.err
  echo "bench tp$N $tag rc=$?"
  tail -1 gpurun_out/bench_tp${N}_$tag.json | head -c 200; echo; tail -2 gpurun_out/bench_tp${N}_$tag.err
}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $((29330 + RANDOM % 200)) tools/ar_bench.py > gpurun_out/ar_bench_n$N.log 2>&1
echo "ar_bench rc=$?"; grep -E "all-reduce|fused|consumer|all-gather|greedy" gpurun_out/ar_bench_n$N.log
run default B200_FUSE_AR_NORM=1
if [ -n "${CFG4_SANITY:-}" ]; then   # config 4's dimensions with 2 layers: exercises hidden 8192 under TP
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port $((29030 + RANDOM % 100)) bench.py --gpus $N --model llama3-70b --layers 2 --batch 32 --seqlen 4096 --quant gptq \
      --steps 3 --warmup 3 --skip-cpu-baseline --no-ttft > gpurun_out/bench_70b_2l_tp$N.json 2> gpurun_out/bench_70b_2l_tp$N.err
  echo "bench 70b dims, 2 layers, tp$N rc=$? $(tail -1 gpurun_out/bench_70b_2l_tp$N.json | head -c 200)"; tail -3 gpurun_out/bench_70b_2l_tp$N.err
fi
if [ -n "${CFG4:-}" ]; then   # Llama-3-70B GPTQ, TP=8, 32 x 4096
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port $((29130 + RANDOM % 200)) bench.py --gpus $N --model llama3-70b --batch 32 --seqlen 4096 --quant gptq \
      --steps 10 --warmup 3 --skip-cpu-baseline --no-ttft > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err
  echo "bench cfg4 rc=$? $(tail -1 gpurun_out/bench_cfg4.json | head -c 300)"; tail -3 gpurun_out/bench_cfg4.err
fi
if [ "${2:-}" != quick ]; then
  run oneshot B200_AR_ALGO=oneshot
  run nccl_gather B200_AR_GATHER=0
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $((29730 + RANDOM % 200)) tools/step_timeline.py --out gpurun_out/step_timeline_tp$N.md \
    > gpurun_out/step_timeline_tp$N.log 2>&1
echo "timeline tp$N rc=$?"; head -60 gpurun_out/step_timeline_tp$N.md 2>/dev/null
if [ "${2:-}" != quick ]; then
  timeout 600 scalellm_b200/decode_demo 32 64 2048 20 $N > gpurun_out/decode_demo_tp$N.log 2>&1
  echo "decode_demo tp$N rc=$? $(tail -1 gpurun_out/decode_demo_tp$N.log)"
fi
