#!/bin/bash
# usage: tools/ncu_step.sh <tag> [kernel-regex-for-full-capture]   (run on the GPU box through gpurun)
# 1. launch list (gpu__time_duration per launch) of OUR kernels over two bench steps;
# 2. ncu --set full (+ source) of the kernels matching the regex, taken from the last step.
TAG=${1:-step}
RE=${2:-"k_scan_pq_db|k_coarse_gemm|k_coarse_select_warp|k_finalize_pq8|k_pq_quantize"}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^k_" -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_${TAG}.log 2>&1
# matching launches: 14 k_pq_quantize while the index is built, then 5 per step (3 warm-up + recall step + ...)
ncu --set full --clock-control none --import-source on -k "regex:${RE}" --launch-skip 29 -c 5 -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
