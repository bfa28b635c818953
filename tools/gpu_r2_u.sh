#!/bin/bash
# round 2: the default kernel across cohort widths (slab ~8 GB each), and an ncu capture of the 1M-sample shape
mkdir -p gpurun_out
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --brief"
for n in 50000 100000 200000 300000 400000 500000 600000 700000 800000 1000000 1180000; do
  v=$(( 4000000000 / n )); v=$(( v / 64 * 64 ))
  echo "n=$n V=$v: $(timeout 200 $B --samples $n --variants $v 2>&1 | tail -1)"
done | tee gpurun_out/width_sweep.txt
echo "== 1M tuning"
for e in "NPC_TILE_A=1" "NPC_TILE_SLEEP=0" "NPC_TILE_SC=10 NPC_TILE_SR=2" ; do
  echo "$e: $(env $e timeout 200 $B --samples 1000000 --variants 4096 2>&1 | tail -1)"
done
echo "== ncu 1M"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused_pair -s 3 -c 1 -o gpurun_out/pair_1M -f $B --samples 1000000 --variants 4096 --steps 2 > gpurun_out/ncu_1M.log 2>&1; tail -2 gpurun_out/ncu_1M.log
