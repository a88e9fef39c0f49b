#!/bin/bash
# ncu capture of the HNSW batch kernel on config 4 (run on the GPU box): the 4th launch is a steady-state batch
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:k_hnsw_spec" -s 3 -c 1 -f -o gpurun_out/prof_c4 \
    python bench.py --config c4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c4.log 2>&1
