#!/bin/bash
# experiment: which of the changes costs the partial-warp shapes 5 %?  A: predication + 768 bound, B: no predication, C: 640 bound, D: neither, E: previous commit
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --brief"
cp nimpress_b200/lib/libnimpress_cuda.so /tmp/keep.so
for v in E A B C D; do
  cp nimpress_b200/lib/variants/$v.so nimpress_b200/lib/libnimpress_cuda.so
  nc1=20; [ $v = C ] && nc1=16; [ $v = D ] && nc1=16
  for nv in "500000 8000" "100000 40000" "800000 4992" "600000 6656"; do
    set -- $nv
    echo "$v n=$1: $(NPC_TILE_NC1=$nc1 timeout 200 $B --samples $1 --variants $2 2>&1 | tail -1 | cut -c1-60)"
  done
done
cp /tmp/keep.so nimpress_b200/lib/libnimpress_cuda.so
