#!/bin/bash
# round 2: raw-ring depth for 17-20 consumer warps; the wide (> 1.2 M) path after the shape changes
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --brief"
for n in 620000 660000 700000 740000; do
  v=$(( 4000000000 / n )); v=$(( v / 64 * 64 ))
  echo "n=$n: $(timeout 200 $B --samples $n --variants $v 2>&1 | tail -1 | cut -c1-330)"
  echo "   SR=3: $(NPC_TILE_SR=3 timeout 200 $B --samples $n --variants $v 2>&1 | tail -1 | cut -c1-330)"
done
echo "2M: $(timeout 300 $B --samples 2000000 --variants 2048 --steps 5 2>&1 | tail -1 | cut -c1-330)"
echo "1.5M: $(timeout 300 $B --samples 1500000 --variants 2560 --steps 5 2>&1 | tail -1 | cut -c1-330)"
