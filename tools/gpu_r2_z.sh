#!/bin/bash
# round 2: four partial warps (one per scheduler) share the slab's end at 1 or 2 words per lane: parity, fuzz, sweep
mkdir -p gpurun_out
echo "== parity"; timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -4
echo "== fuzz"; timeout 400 python tools/fuzz_parity.py --cases 600 --seed 35 --seconds 200 2>&1 | tail -2
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --brief"
for n in 50000 100000 250000 400000 500000 700000 800000; do
  v=$(( 4000000000 / n )); v=$(( v / 64 * 64 ))
  echo "n=$n V=$v: $(timeout 200 $B --samples $n --variants $v 2>&1 | tail -1)"
  echo "   TAIL=0: $(NPC_TILE_TAIL=0 timeout 200 $B --samples $n --variants $v 2>&1 | tail -1 | cut -c1-75)"
done | tee gpurun_out/tail_sweep2.txt
echo "== exact"; NPC_EXACT=1 timeout 200 $B --samples 500000 --variants 8000 2>&1 | tail -1 | cut -c1-75
echo "== config 2 / 5"; timeout 300 python bench.py --config 2 2>&1 | tail -1 | cut -c1-300; timeout 300 python bench.py --config 5 2>&1 | tail -1 | cut -c1-300
