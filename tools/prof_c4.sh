#!/bin/bash
# ncu capture of the HNSW kernel on config 4 (run on the GPU box): launch 6 = a steady-state batch
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:k_hnsw_search_reg" --launch-skip 5 -c 1 -f -o gpurun_out/prof_c4 \
    python tools/bench_configs.py c4 > gpurun_out/ncu_c4.log 2>&1
