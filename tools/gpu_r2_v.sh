#!/bin/bash
# round 2: chunk-set raw stages + predicated idle lanes + 20 consumer warps at K = 1 + 4-tile decider groups: parity, then the width sweep
mkdir -p gpurun_out
echo "== parity"; timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -4
echo "== fuzz"; timeout 300 python tools/fuzz_parity.py --cases 300 --seed 31 --seconds 120 2>&1 | tail -2
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --brief"
for n in 50000 100000 200000 300000 400000 500000 600000 700000 800000 1000000 1180000; do
  v=$(( 4000000000 / n )); v=$(( v / 64 * 64 ))
  echo "n=$n V=$v: $(timeout 200 $B --samples $n --variants $v 2>&1 | tail -1)"
done | tee gpurun_out/width_sweep2.txt
echo "== tuning"
echo "400k GR=1: $(NPC_TILE_GR=1 timeout 200 $B --samples 400000 --variants 9984 2>&1 | tail -1)"
echo "800k SR=5: $(NPC_TILE_SR=5 timeout 200 $B --samples 800000 --variants 4992 2>&1 | tail -1)"
echo "800k SR=3: $(NPC_TILE_SR=3 timeout 200 $B --samples 800000 --variants 4992 2>&1 | tail -1)"
echo "1M SR=3: $(NPC_TILE_SR=3 timeout 200 $B --samples 1000000 --variants 3968 2>&1 | tail -1)"
echo "1M GD=8: $(NPC_TILE_GD=8 timeout 200 $B --samples 1000000 --variants 3968 2>&1 | tail -1)"
echo "700k NC1=16 (K=2): $(NPC_TILE_NC1=16 timeout 200 $B --samples 700000 --variants 5696 2>&1 | tail -1)"
echo "500k GD=4: $(NPC_TILE_GD=4 timeout 200 $B --samples 500000 --variants 8000 2>&1 | tail -1)"
echo "600k SR=4: $(NPC_TILE_SR=4 timeout 200 $B --samples 600000 --variants 6656 2>&1 | tail -1)"
echo "== int16"; timeout 300 python tools/bench_int16.py 2>&1 | tail -1
