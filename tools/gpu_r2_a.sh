#!/bin/bash
# round 2, batch A: parity of the pair-lookup kernel (v5) + A/B against the round-1 kernel (v4) + ncu of v5
mkdir -p gpurun_out
echo "== parity (v5 default)"; timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -4
B="python bench.py --variants 32768 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --brief"
echo "== v4"; NPC_TILE_V=4 timeout 200 $B 2>&1 | tail -1
echo "== v5"; timeout 200 $B 2>&1 | tail -1
echo "== v5 sleep 200"; NPC_TILE_SLEEP=200 timeout 200 $B 2>&1 | tail -1
echo "== v5 A=1"; NPC_TILE_A=1 timeout 200 $B 2>&1 | tail -1
echo "== v5 SC=16"; NPC_TILE_SC=16 timeout 200 $B 2>&1 | tail -1
echo "== v5 SR=3"; NPC_TILE_SR=3 timeout 200 $B 2>&1 | tail -1
echo "== v5 SR=5 SC=14"; NPC_TILE_SR=5 NPC_TILE_SC=14 timeout 200 $B 2>&1 | tail -1
echo "== v5 exact"; NPC_EXACT=1 timeout 200 $B 2>&1 | tail -1
echo "== v4 exact"; NPC_EXACT=1 NPC_TILE_V=4 timeout 200 $B 2>&1 | tail -1
echo "== shapes v5"; bash tools/gpu_shapes.sh
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_fused_pair' -s 1 -c 1 -o gpurun_out/prof_r2_pair_a -f \
    python bench.py --variants 32768 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_r2_a.log 2>&1; tail -1 gpurun_out/ncu_r2_a.log | cut -c1-200
