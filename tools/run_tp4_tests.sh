#!/bin/bash
# 4-GPU stage: the all-reduce suite at 4 and 2 ranks (every part of the worker), nothing else.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=tests/test_gpu_allreduce.py
timeout 400 python -m pytest "$T::test_nvlink_allreduce_matches_nccl_and_host_sum[4-]" \
    "$T::test_nvlink_allreduce_matches_nccl_and_host_sum[2-twoshot]" -m gpu -q --tb=short -p no:cacheprovider \
    > gpurun_out/pytest_allreduce_n4.log 2>&1
echo "pytest allreduce (4 GPUs) rc=$? : $(tail -1 gpurun_out/pytest_allreduce_n4.log)"
grep -E "Error|assert|FAILED" gpurun_out/pytest_allreduce_n4.log | head -10 | cut -c1-300
