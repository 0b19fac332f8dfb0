cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_allreduce.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_allreduce.log 2>&1
echo "pytest allreduce rc=$? : $(tail -3 gpurun_out/pytest_allreduce.log | tr '\n' ' ')"
grep -E "Error|assert" gpurun_out/pytest_allreduce.log | head -10
for f in 1 0; do
B200_FUSE_AR_NORM=$f timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$f bench.py --gpus $N --steps 20 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_tp${N}_f$f.json 2> gpurun_out/bench_tp${N}_f$f.err
echo "bench tp$N fuse_ar_norm=$f rc=$?"
tail -1 gpurun_out/bench_tp${N}_f$f.json | head -c 200; echo; tail -2 gpurun_out/bench_tp${N}_f$f.err
done
