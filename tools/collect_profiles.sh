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
ls -la $P
