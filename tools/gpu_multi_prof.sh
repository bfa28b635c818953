#!/bin/bash
# ncu capture of the multi-score contraction kernel (dense shape: 20000 variants x 200k samples x 18 definitions)
mkdir -p gpurun_out
NPC_MULTI=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_multi_contract' -s 1 -c 1 -o gpurun_out/prof_multi -f \
    python tools/bench_multi.py --variants ${1:-20000} --reps 1 > gpurun_out/ncu_multi.log 2>&1; tail -2 gpurun_out/ncu_multi.log | cut -c1-300
