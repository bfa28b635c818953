#!/bin/bash
# round 2: K = 1 on up to 24 warps (72-register instance above 16), K = 2 with one raw stage per chunk set: parity, sweep
mkdir -p gpurun_out
echo "== parity"; timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -4
echo "== fuzz"; timeout 300 python tools/fuzz_parity.py --cases 300 --seed 33 --seconds 100 2>&1 | tail -2
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --brief"
for n in 50000 100000 150000 200000 300000 400000 500000 600000 700000 800000 900000 1000000 1100000 1180000; do
  v=$(( 4000000000 / n )); v=$(( v / 64 * 64 ))
  echo "n=$n V=$v: $(timeout 200 $B --samples $n --variants $v 2>&1 | tail -1)"
done | tee gpurun_out/width_sweep4.txt
echo "== tuning"
echo "400k GR=1: $(NPC_TILE_GR=1 timeout 200 $B --samples 400000 --variants 9984 2>&1 | tail -1)"
echo "500k GR=2: $(NPC_TILE_GR=2 timeout 200 $B --samples 500000 --variants 8000 2>&1 | tail -1)"
echo "250k GR=1: $(NPC_TILE_GR=1 timeout 200 $B --samples 250000 --variants 16000 2>&1 | tail -1)"
echo "250k: $(timeout 200 $B --samples 250000 --variants 16000 2>&1 | tail -1)"
echo "900k GD=8: $(NPC_TILE_GD=8 timeout 200 $B --samples 900000 --variants 4416 2>&1 | tail -1)"
echo "900k SR=3: $(NPC_TILE_SR=3 timeout 200 $B --samples 900000 --variants 4416 2>&1 | tail -1)"
echo "== exact"; NPC_EXACT=1 timeout 200 $B --samples 500000 --variants 8000 2>&1 | tail -1
NPC_EXACT=1 timeout 200 $B --samples 850000 --variants 4672 2>&1 | tail -1
