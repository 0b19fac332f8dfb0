cd /root/repo
mkdir -p gpurun_out
timeout 120 tools/microbench/umma > gpurun_out/micro_umma.log 2>&1; echo "umma rc=$?"; cat gpurun_out/micro_umma.log
for f in gpu_w4a16 gpu_decode_step; do
  timeout 900 python -m pytest tests/test_$f.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_$f.log 2>&1
  echo "pytest $f rc=$? : $(tail -1 gpurun_out/pytest_$f.log)"
done
timeout 300 python tools/w4_trace.py > gpurun_out/w4_trace.log 2>&1; echo "trace rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 --skip-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
B200_W4_NSUB=2 timeout 600 python bench.py --steps 20 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_nsub2.json 2> gpurun_out/bench_nsub2.err; echo "bench nsub2 rc=$?"
for f in bench bench_nsub2; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1])
    g=d["roofline_w4a16_gemm"]["per_proj"]
    print("$f", round(d["value"]), "tok/s", round(d["ms_per_step"],3), "ms; e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"], "attn us", round(d["roofline"]["us_per_launch"],1), "gemm us", {k: round(v["us"],1) for k,v in g.items()})
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/$f.err").read()[-1500:])
PY
done
