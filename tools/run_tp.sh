cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 600 python -m pytest tests/test_gpu_allreduce.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_allreduce.log 2>&1
echo "pytest allreduce rc=$? : $(tail -3 gpurun_out/pytest_allreduce.log | tr '\n' ' ')"
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_tp$N.json 2> gpurun_out/bench_tp$N.err
echo "bench tp$N rc=$?"
tail -c 1800 gpurun_out/bench_tp$N.json; tail -5 gpurun_out/bench_tp$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_tp$N.json 2> gpurun_out/bench_ref_tp$N.err
echo "bench ref tp$N rc=$?"; tail -c 600 gpurun_out/bench_ref_tp$N.json
