#!/bin/bash
# round 2: where is the crossover between the short-launch and the long-launch grid split?
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --brief"
for nv in "100000 697" "200000 730" "250000 512" "250000 2048" "500000 128" "500000 256" "500000 512" "500000 1024" "500000 2048"; do
  set -- $nv
  echo "n=$1 V=$2 ($(( $1 * $2 * 2 / 1000000 )) MB): short $(NPC_TILE_LONG_MB=100000 timeout 200 $B --samples $1 --variants $2 2>&1 | tail -1 | cut -c1-62)"
  echo "                                   long  $(NPC_TILE_LONG_MB=1 timeout 200 $B --samples $1 --variants $2 2>&1 | tail -1 | cut -c1-62)"
done
