cd /root/repo
mkdir -p gpurun_out
timeout 120 tools/microbench/mix > gpurun_out/micro_mix.log 2>&1; cat gpurun_out/micro_mix.log
for f in gpu_w4a16 gpu_decode_step; do
timeout 900 python -m pytest tests/test_$f.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_$f.log 2>&1
echo "pytest $f rc=$? : $(tail -1 gpurun_out/pytest_$f.log)"
done
for ns in 1 2; do
B200_W4_NSUB=$ns timeout 300 python tools/w4_trace.py > gpurun_out/w4_trace_ns$ns.log 2>&1; echo "trace ns$ns rc=$?"
B200_W4_NSUB=$ns timeout 600 python bench.py --steps 20 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_ns$ns.json 2> gpurun_out/bench_ns$ns.err; echo "bench rc=$?"
python - <<PY
import json
f="bench_ns$ns"
try:
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    g=d["roofline_w4a16_gemm"]["per_proj"]
    print(f, round(d["value"]), "tok/s", round(d["ms_per_step"],3), "ms; e2e", round(d["e2e"]["value"]), "gemm us", {k: round(v["us"],1) for k,v in g.items()})
except Exception as e:
    print(f, "failed", e); print(open("gpurun_out/%s.err"%f).read()[-1500:])
PY
done
