#!/bin/bash
# round 2, batch G: the ncu captures VERDICT r1 asked for -- exact-order mode, the config-2 launch shape, the generic (two-kernel) path
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
NPC_EXACT=1 timeout 500 $NCU -k regex:'k_fused_pair' -s 1 -c 1 -o gpurun_out/prof_r2_pair_exact -f \
    python bench.py --variants 32768 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extra > gpurun_out/ncu_g1.log 2>&1; tail -1 gpurun_out/ncu_g1.log | cut -c1-160
timeout 500 $NCU -k regex:'k_fused_pair|k_add_partials' -s 4 -c 2 -o gpurun_out/prof_r2_pair_config2 -f \
    python bench.py --config 2 --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --no-extra > gpurun_out/ncu_g2.log 2>&1; tail -1 gpurun_out/ncu_g2.log | cut -c1-160
echo "== int32 generic"; timeout 300 python tools/bench_int16.py --width 4 --variants 2048 2>&1 | tail -1
timeout 500 $NCU -k regex:'k_count_generic|k_accum_generic|k_decide' -s 3 -c 3 -o gpurun_out/prof_r2_generic_int32 -f \
    python tools/bench_int16.py --width 4 --variants 2048 --steps 1 > gpurun_out/ncu_g3.log 2>&1; tail -1 gpurun_out/ncu_g3.log | cut -c1-160
echo "== 2M samples two-kernel"; timeout 500 $NCU -k regex:'k_count_i8x2|k_accum_i8x2' -s 2 -c 2 -o gpurun_out/prof_r2_two_kernel_2M -f \
    python bench.py --samples 2000000 --variants 2048 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extra > gpurun_out/ncu_g4.log 2>&1; tail -1 gpurun_out/ncu_g4.log | cut -c1-160
