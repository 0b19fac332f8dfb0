#!/bin/bash
# Turn the scratch output of `tools/gpu_ci.sh bench ncu micro trace` (gpurun_out/) into the committed
# summaries under profiles/.   usage: tools/collect_profiles.sh r01
set -u
cd "$(dirname "$0")/.."
R=${1:-r01}
G=gpurun_out
P=profiles
mkdir -p $P
[ -s $G/bench.json ] && tail -1 $G/bench.json > $P/${R}_bench.json
[ -s $G/bench_ref.json ] && tail -1 $G/bench_ref.json > $P/${R}_bench_reference.json
[ -s $G/launches.csv ] && python tools/ncu_summary.py launches $G/launches.csv $P/${R}_launches_4layers_eager.md \
    "$R: launch list of bench.py --layers 4 --no-graph (1 warm-up + 1 step)" > /dev/null
[ -s $G/prof_attn.ncu-rep ] && python tools/ncu_summary.py full $G/prof_attn.ncu-rep $P/${R}_ncu_paged_attn.md paged_attn > /dev/null
[ -s $G/prof_gemm.ncu-rep ] && python tools/ncu_summary.py full $G/prof_gemm.ncu-rep $P/${R}_ncu_w4a16_gemm.md w4a16_gemm > /dev/null
if ls $G/micro_*.log > /dev/null 2>&1; then
  { echo "# $R: microbenchmarks on the B200 box (tools/microbench/*.cu)"; echo;
    for f in $G/micro_*.log; do echo "## $(basename $f .log)"; echo '```'; cat $f; echo '```'; echo; done; } > $P/${R}_microbench.md
fi
[ -s $G/w4_trace.log ] && { echo "# $R: device-side clock64 trace of w4a16_gemm_kernel (tools/w4_trace.py)"; echo '```'; cat $G/w4_trace.log; echo '```'; } > $P/${R}_w4_trace.md
[ -s $G/attn_bench.log ] && { echo "# $R: attention kernel-only timings (tools/attn_bench.py)"; echo '```'; cat $G/attn_bench.log; echo '```'; } > $P/${R}_attn_bench.md
[ -s $G/attn_variants.log ] && { echo "# $R: kernel-only A/B of the attention stream kernel's instantiations (tools/attn_bench.py variant)"; echo '```'; grep "^attn\|^---" $G/attn_variants.log; echo '```'; } > $P/${R}_attn_variants.md
[ -s $G/step_timeline.md ] && cp $G/step_timeline.md $P/${R}_step_timeline.md
[ -s $G/w4_variants.jsonl ] && { echo "# $R: GEMM-only A/B of the W4A16 kernel variants (tools/w4_variant_bench.py; us per launch, CUDA graph replay, M = 64)"; echo '```'; cat $G/w4_variants.jsonl; echo '```'; } > $P/${R}_w4_variants.md
# A/B runs of the opt-in variants: one line per run (value, ms/step, GEMM launch times)
if ls $G/bench_w4var*.json $G/bench_occ*.json 2> /dev/null | grep -q .; then
  python - $G $P/${R}_variants.md "$R" <<'PY'
import glob, json, os, sys
g, out, r = sys.argv[1:4]
rows = []
for f in sorted(glob.glob(os.path.join(g, "bench_w4var*.json")) + glob.glob(os.path.join(g, "bench_occ*.json"))):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception:
        continue
    gm = (d.get("roofline_w4a16_gemm") or {}).get("per_proj") or {}
    rows.append((os.path.basename(f), d["value"], d["ms_per_step"], (d.get("roofline") or {}).get("us_per_launch"),
                 " / ".join(f"{v['us']:.1f}" for v in gm.values())))
with open(out, "w") as fh:
    fh.write(f"# {r}: A/B bench runs of the opt-in kernel variants (same box, back to back)\n\n"
             "| run | tokens/s | ms/step | attention us/launch | GEMM us (qkv / o / gate_up / down) |\n|---|---:|---:|---:|---|\n")
    for n, v, ms, au, gm in rows:
        fh.write(f"| `{n}` | {v:.0f} | {ms:.3f} | {au if au is None else round(au, 1)} | {gm} |\n")
PY
fi
ls -la $P
