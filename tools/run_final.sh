#!/bin/bash
# Final single-GPU evidence run of a round (gpurun -- bash tools/run_final.sh): full GPU test suite, smoke, the bench
# line and the reference arm, the ncu launch list + full captures of the hot kernels, the in-graph timeline.
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --tb=short -p no:cacheprovider > $OUT/pytest_all.log 2>&1
echo "pytest rc=$? : $(tail -1 $OUT/pytest_all.log)" | tee $OUT/summary.txt
grep -E "^(FAILED|ERROR)|Error|^E  " $OUT/pytest_all.log | head -20 | cut -c1-300 | tee -a $OUT/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1
echo "smoke rc=$? $(tail -1 $OUT/smoke.log)" | tee -a $OUT/summary.txt
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$? $(tail -1 $OUT/bench.json | head -c 250)" | tee -a $OUT/summary.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
echo "bench reference arm rc=$? $(tail -1 $OUT/bench_ref.json | head -c 300)" | tee -a $OUT/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 1 --layers 4 --no-graph --no-ttft \
    --skip-cpu-baseline > $OUT/ncu_bench.log 2>&1
echo "ncu launches rc=$?" | tee -a $OUT/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:paged_attn_persist_kernel -s 4 -c 2 -o $OUT/prof_attn -f \
    python bench.py --steps 1 --warmup 1 --layers 4 --no-graph --no-ttft --skip-cpu-baseline > $OUT/ncu_attn.log 2>&1
echo "ncu attn rc=$?" | tee -a $OUT/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:w4a16_gemm_kernel -s 8 -c 4 -o $OUT/prof_gemm -f \
    python bench.py --steps 1 --warmup 1 --layers 4 --no-graph --no-ttft --skip-cpu-baseline > $OUT/ncu_gemm.log 2>&1
echo "ncu gemm rc=$?" | tee -a $OUT/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:dense_gemm_kernel -c 1 -o $OUT/prof_dense -f \
    python bench.py --steps 1 --warmup 1 --layers 2 --no-graph --no-ttft --skip-cpu-baseline > $OUT/ncu_dense.log 2>&1
echo "ncu dense rc=$?" | tee -a $OUT/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prefill_attn -s 3 -c 1 -f \
    -o $OUT/prof_prefill python tools/prefill_trace.py > $OUT/ncu_prefill.log 2>&1
echo "ncu prefill rc=$?" | tee -a $OUT/summary.txt
timeout 600 python tools/step_timeline.py --out $OUT/step_timeline.md > $OUT/step_timeline.log 2>&1
echo "timeline rc=$?" | tee -a $OUT/summary.txt
