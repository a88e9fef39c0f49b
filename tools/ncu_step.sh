#!/bin/bash
# usage: tools/ncu_step.sh <tag> [kernel-regex-for-full-capture]   (run on the GPU box through gpurun)
# 1. launch list (gpu__time_duration per launch) of OUR kernels over two bench steps (c3);
# 2. ncu --set full (+ source) of one launch of the scan kernel and of the kernels matching the regex.
# Summaries: python profiles/summarize.py <tag> gpurun_out/launches_<tag>.csv [gpurun_out/prof_<tag>_scan.ncu-rep <regex>]
TAG=${1:-step}
RE=${2:-"k_finalize|k_coarse_gemm|k_coarse_select_q|k_pq_quantize"}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^k_" -c 120 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_scan_pq16" -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_scan \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_scan_${TAG}.log 2>&1
ncu --set full --clock-control none -k "regex:${RE}" -s 12 -c 4 -f -o gpurun_out/prof_${TAG}_rest \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_rest_${TAG}.log 2>&1
