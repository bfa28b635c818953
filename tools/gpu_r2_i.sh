#!/bin/bash
mkdir -p gpurun_out
echo "== wide tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "wider or full_sample or int16" 2>&1 | tail -3
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --brief"
echo "== 2M samples x 2048 (decided-mode slabs)"; timeout 120 $B --samples 2000000 --variants 2048 2>&1 | tail -1
echo "== 2M exact"; NPC_EXACT=1 timeout 120 $B --samples 2000000 --variants 2048 2>&1 | tail -1
echo "== bench_multi"; NPC_TIMING=1 timeout 600 python tools/bench_multi.py --reps 5 > gpurun_out/bench_multi_r2b.json 2> gpurun_out/bench_multi_r2b.err; cat gpurun_out/bench_multi_r2b.json | cut -c1-400; grep "npc multi" gpurun_out/bench_multi_r2b.err | tail -7
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_multi_contract' -s 2 -c 1 -o gpurun_out/prof_r2_multi_contract -f \
    python tools/bench_multi.py --variants 20000 --reps 2 > gpurun_out/ncu_multi_r2.log 2>&1; tail -1 gpurun_out/ncu_multi_r2.log | cut -c1-200
