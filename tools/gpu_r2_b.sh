#!/bin/bash
# round 2, batch B: all GPU tests with the dp4a pair lookup, sweeps, ncu, the full default bench line
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
B="python bench.py --variants 32768 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --brief"
echo "== v5 dp4a"; timeout 200 $B 2>&1 | tail -1
echo "== v5 SR=3"; NPC_TILE_SR=3 timeout 200 $B 2>&1 | tail -1
echo "== v5 SR=3 A=1"; NPC_TILE_SR=3 NPC_TILE_A=1 timeout 200 $B 2>&1 | tail -1
echo "== v5 SR=2"; NPC_TILE_SR=2 timeout 200 $B 2>&1 | tail -1
echo "== v5 exact"; NPC_EXACT=1 timeout 200 $B 2>&1 | tail -1
echo "== shapes"; bash tools/gpu_shapes.sh
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_fused_pair' -s 1 -c 1 -o gpurun_out/prof_r2_pair_b -f \
    python bench.py --variants 32768 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extra > gpurun_out/ncu_r2_b.log 2>&1; tail -1 gpurun_out/ncu_r2_b.log | cut -c1-200
echo "== full bench"; timeout 1200 python bench.py 2>&1 | tail -1 > gpurun_out/bench_r2_b.json; cut -c1-1500 gpurun_out/bench_r2_b.json
