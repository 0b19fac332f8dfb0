#!/bin/bash
# Runs on the B200 box (via gpurun).  Every stage is its own process under `timeout` so a hung
# kernel cannot wedge the whole call; logs land in gpurun_out/.
#   tools/gpu_ci.sh [tests] [probe] [smoke] [bench] [ncu]     (default: all of them)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
STAGES="${*:-tests probe smoke bench ncu}"
# "round2" = everything that was staged after round 1's GPU budget ran out, in one call
if [[ " $STAGES " == *" round2 "* ]]; then
  STAGES="$STAGES tests smoke bench staged timeline trace micro occ refk"
fi
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
echo "stages: $STAGES" | tee $OUT/summary.txt
has() { [[ " $STAGES " == *" $1 "* ]]; }

if has probe; then
  for impl in simt tcgen05; do
    for t in t1 t2 t3 t4 t5 t6; do
      if [ $impl = simt ] && [ $t != t2 ] && [ $t != t4 ]; then continue; fi
      B200_W4A16_IMPL=$impl timeout 120 python tools/w4_probe.py $t >> $OUT/probe.log 2>&1
      echo "probe $impl $t rc=$?" >> $OUT/summary.txt
    done
  done
  grep -h "^\[" $OUT/probe.log | tee -a $OUT/summary.txt
fi

if has trace; then
  timeout 300 python tools/w4_trace.py > $OUT/w4_trace.log 2>&1
  echo "trace rc=$?" | tee -a $OUT/summary.txt
  B200_PDL=0 timeout 300 python tools/w4_trace.py >> $OUT/w4_trace.log 2>&1
  echo "trace (B200_PDL=0) rc=$?" | tee -a $OUT/summary.txt
  cat $OUT/w4_trace.log >> $OUT/summary.txt
fi

if has attn2; then
  # both attention variants through the same parity tests + a kernel-only timing
  for impl in mma simt; do
    B200_ATTN_IMPL=$impl timeout 900 python -m pytest tests/test_gpu_attention.py -m gpu -q --tb=short \
        -p no:cacheprovider > $OUT/pytest_attention_$impl.log 2>&1
    echo "pytest attention[$impl] rc=$? : $(tail -1 $OUT/pytest_attention_$impl.log)" | tee -a $OUT/summary.txt
    B200_ATTN_IMPL=$impl timeout 600 python tools/attn_bench.py >> $OUT/attn_bench.log 2>&1
  done
  cat $OUT/attn_bench.log | tee -a $OUT/summary.txt
fi

if has tests; then
  for f in gpu_elementwise gpu_attention gpu_w4a16 gpu_decode_step shim cpp_host; do
    timeout 1200 python -m pytest tests/test_$f.py -m gpu -q --tb=short -p no:cacheprovider \
        > $OUT/pytest_$f.log 2>&1
    echo "pytest $f rc=$? : $(tail -1 $OUT/pytest_$f.log)" | tee -a $OUT/summary.txt
  done
fi

if has smoke; then
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1
  echo "smoke rc=$? : $(tail -1 $OUT/smoke.log)" | tee -a $OUT/summary.txt
fi

if has bench; then
  timeout 1500 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
  echo "bench rc=$?" | tee -a $OUT/summary.txt
  tail -c 3000 $OUT/bench.json | tee -a $OUT/summary.txt
  tail -5 $OUT/bench.err >> $OUT/summary.txt
fi

if has ncu; then
  # launch list of a short eager run (cold-cache, serialised: compare SHARES only)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
      --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 1 --layers 4 --no-graph \
      --skip-cpu-baseline > $OUT/ncu_bench.log 2>&1
  echo "ncu launches rc=$?" | tee -a $OUT/summary.txt
  # full capture of the dominant kernel
  timeout 900 ncu --set full --clock-control none --import-source on \
      -k regex:paged_attn_persist_kernel -s 4 -c 2 -o $OUT/prof_attn -f \
      python bench.py --steps 1 --warmup 1 --layers 4 --no-graph --skip-cpu-baseline \
      > $OUT/ncu_attn.log 2>&1
  echo "ncu attn rc=$?" | tee -a $OUT/summary.txt
  timeout 900 ncu --set full --clock-control none --import-source on \
      -k regex:w4a16_gemm_kernel -s 8 -c 4 -o $OUT/prof_gemm -f \
      python bench.py --steps 1 --warmup 1 --layers 4 --no-graph --skip-cpu-baseline \
      > $OUT/ncu_gemm.log 2>&1
  echo "ncu gemm rc=$?" | tee -a $OUT/summary.txt
fi
if has ncul; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 400 --csv \
      --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 1 --layers 8 --no-graph \
      --skip-cpu-baseline > $OUT/ncu_bench.log 2>&1
  echo "ncu launches rc=$?" | tee -a $OUT/summary.txt
fi

if has ncuattn; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:paged_attn_mma_kernel \
      -s 6 -c 2 -o $OUT/prof_attn_mma -f python tools/attn_bench.py > $OUT/ncu_attn_mma.log 2>&1
  echo "ncu attn mma rc=$?" | tee -a $OUT/summary.txt
fi
echo "== done" | tee -a $OUT/summary.txt
if has deq; then
  timeout 120 tools/microbench/deq > $OUT/deq.log 2>&1
  echo "deq rc=$?" | tee -a $OUT/summary.txt
  cat $OUT/deq.log | tee -a $OUT/summary.txt
fi
if has refk; then
  # our kernels against the reference's own kernels (oracle/_ref), then the attention A/B timing
  timeout 900 python -m pytest tests/test_gpu_vs_reference_kernels.py -m gpu -q \
      --tb=short -p no:cacheprovider > $OUT/pytest_vs_reference.log 2>&1
  echo "pytest vs reference kernels rc=$? : $(tail -1 $OUT/pytest_vs_reference.log)" | tee -a $OUT/summary.txt
  timeout 600 python tools/attn_bench.py > $OUT/attn_bench.log 2>&1
  cat $OUT/attn_bench.log | tee -a $OUT/summary.txt
  timeout 600 python tools/w4_ref_bench.py > $OUT/w4_ref_bench.log 2>&1
  cat $OUT/w4_ref_bench.log | tee -a $OUT/summary.txt
fi
if has occ; then
  # opt-in instantiations of the attention stream kernel: B200_ATTN_OCC=1 (2 stages, 11 CTAs/SM),
  # B200_ATTN_TR=1 (transposed tile), and both: parity (incl. the staged <= 8-row variants test),
  # kernel-only timing, then the full bench for the default and the fastest
  : > $OUT/attn_variants.log
  for v in "0 0" "1 0" "0 1" "1 1"; do
    set -- $v
    if [ "$v" != "0 0" ]; then
      B200_ATTN_OCC=$1 B200_ATTN_TR=$2 timeout 900 python -m pytest tests/test_gpu_attention.py \
          tests/test_gpu_decode_step.py -m gpu -q --tb=short -p no:cacheprovider > $OUT/pytest_attention_occ$1_tr$2.log 2>&1
      echo "pytest attention[occ=$1 tr=$2] rc=$? : $(tail -1 $OUT/pytest_attention_occ$1_tr$2.log)" | tee -a $OUT/summary.txt
    fi
    B200_ATTN_OCC=$1 B200_ATTN_TR=$2 timeout 300 python tools/attn_bench.py variant >> $OUT/attn_variants.log 2>&1
  done
  grep "^attn" $OUT/attn_variants.log | tee -a $OUT/summary.txt
  BEST=$(grep "^attn" $OUT/attn_variants.log | sort -t: -k2 -n | head -1 | sed 's/.*occ=\([01]\) tr=\([01]\).*/\1 \2/')
  for v in "0 0" "$BEST"; do
    set -- $v
    [ "$v" = "0 0" ] && [ "$BEST" = "0 0" ] && [ -s $OUT/bench_occ0_tr0.json ] && continue
    B200_ATTN_OCC=$1 B200_ATTN_TR=$2 timeout 600 python bench.py --steps 20 --warmup 3 --skip-cpu-baseline \
        > $OUT/bench_occ$1_tr$2.json 2> $OUT/bench_occ$1_tr$2.err
    echo "bench occ=$1 tr=$2 rc=$? $(tail -1 $OUT/bench_occ$1_tr$2.json | head -c 200)" | tee -a $OUT/summary.txt
  done
fi
if has pdl; then
  # programmatic-dependent-launch knobs against the default (B200_PDL=1): off, everything, and the
  # attention stream kernel + combine pass launched programmatically too
  for cfg in "B200_PDL=0" "B200_PDL=2" "B200_ATTN_PDL=1"; do
    env $cfg timeout 600 python bench.py --steps 20 --warmup 3 --skip-cpu-baseline \
        > $OUT/bench_pdl_$cfg.json 2> $OUT/bench_pdl_$cfg.err
    echo "bench $cfg rc=$? $(tail -1 $OUT/bench_pdl_$cfg.json | head -c 160)" | tee -a $OUT/summary.txt
  done
fi
if has staged; then
  # the C++ host classes, the prefill path of the int4 linear, TTFT by chunk size, the C++-only demo
  timeout 900 python -m pytest tests/test_cpp_host.py tests/test_gpu_w4a16.py -m gpu -q --tb=short \
      -p no:cacheprovider -k "cuda_graph_step or model_runner or dense_prefill or cpp_decode_step" > $OUT/pytest_staged.log 2>&1
  echo "pytest staged (C++ CudaGraphStep / ModelRunner, dense prefill linear) rc=$? : $(tail -1 $OUT/pytest_staged.log)" | tee -a $OUT/summary.txt
  timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --ttft > $OUT/bench_ttft.json 2> $OUT/bench_ttft.err
  echo "bench --ttft rc=$? $(tail -1 $OUT/bench_ttft.json | python -c 'import json,sys; print(json.loads(sys.stdin.read())["config"]["ttft"])' 2>&1 | head -c 300)" | tee -a $OUT/summary.txt
  timeout 600 scalellm_b200/decode_demo 32 64 2048 20 > $OUT/decode_demo.log 2>&1   # the step driven from C++ only
  echo "decode_demo rc=$? $(tail -1 $OUT/decode_demo.log)" | tee -a $OUT/summary.txt
  for ch in 512 2048; do   # bigger prefill chunks with the int4 linears as dequant + library bf16 GEMM
    B200_W4_PREFILL_DENSE=1 timeout 600 python bench.py --steps 5 --warmup 3 --skip-cpu-baseline --ttft --ttft-chunk $ch \
        > $OUT/bench_ttft_dense$ch.json 2> $OUT/bench_ttft_dense$ch.err
    echo "bench --ttft dense chunk=$ch rc=$? $(tail -1 $OUT/bench_ttft_dense$ch.json | python -c 'import json,sys; print(json.loads(sys.stdin.read())["config"]["ttft"])' 2>&1 | head -c 300)" | tee -a $OUT/summary.txt
  done
fi
if has timeline; then
  # in-graph device timeline of the decode step (CUPTI via torch.profiler): shares, gaps, overlap
  timeout 900 python tools/step_timeline.py > $OUT/step_timeline.log 2>&1
  echo "timeline rc=$?" | tee -a $OUT/summary.txt
  head -40 $OUT/step_timeline.log >> $OUT/summary.txt
fi
if has ref; then
  timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
  echo "bench reference rc=$?" | tee -a $OUT/summary.txt
  tail -c 900 $OUT/bench_ref.json | tee -a $OUT/summary.txt
fi
if has micro; then
  for b in readbw kvbw hmma deq umma mix; do
    timeout 120 tools/microbench/$b > $OUT/micro_$b.log 2>&1
    echo "micro $b rc=$?" | tee -a $OUT/summary.txt
  done
fi
