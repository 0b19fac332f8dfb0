cd /root/repo
mkdir -p gpurun_out
for v in 0 1; do
B200_ATTN_PDL=$v timeout 600 python bench.py --steps 30 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_apdl$v.json 2> gpurun_out/bench_apdl$v.err; echo "bench attn_pdl=$v rc=$?"
python - <<PY
import json
f="bench_apdl$v"
try:
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    g=d["roofline_w4a16_gemm"]["per_proj"]
    print(f, round(d["value"]), "tok/s", round(d["ms_per_step"],3), "ms; e2e", round(d["e2e"]["value"]), "attn us", round(d["roofline"]["us_per_launch"],1), "gemm us", {k: round(v["us"],1) for k,v in g.items()})
except Exception as e:
    print(f, "failed", e); print(open("gpurun_out/%s.err"%f).read()[-1500:])
PY
done
