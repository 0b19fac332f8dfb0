cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L | head -8
if [ "${2:-}" = "test" ]; then
timeout 600 python -m pytest tests/test_gpu_allreduce.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_allreduce.log 2>&1
echo "pytest allreduce rc=$? : $(tail -3 gpurun_out/pytest_allreduce.log | tr '\n' ' ')"
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_tp$N.json 2> gpurun_out/bench_tp$N.err
echo "bench tp$N rc=$?"
tail -1 gpurun_out/bench_tp$N.json | head -c 330; echo; tail -3 gpurun_out/bench_tp$N.err
