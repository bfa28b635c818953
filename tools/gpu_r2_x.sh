#!/bin/bash
# round 2: joint / split raw stages, 4-tile decider groups, 20 (24) consumer warps at K = 1: parity, then the width sweep
mkdir -p gpurun_out
echo "== parity"; timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -4
echo "== fuzz"; timeout 300 python tools/fuzz_parity.py --cases 300 --seed 32 --seconds 100 2>&1 | tail -2
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --brief"
for n in 50000 100000 200000 300000 400000 500000 600000 700000 800000 900000 1000000 1180000; do
  v=$(( 4000000000 / n )); v=$(( v / 64 * 64 ))
  echo "n=$n V=$v: $(timeout 200 $B --samples $n --variants $v 2>&1 | tail -1)"
done | tee gpurun_out/width_sweep3.txt
echo "== tuning"
echo "1M split: $(NPC_TILE_SPLIT=1 timeout 200 $B --samples 1000000 --variants 3968 2>&1 | tail -1)"
echo "1.18M joint: $(NPC_TILE_SPLIT=0 timeout 200 $B --samples 1180000 --variants 3328 2>&1 | tail -1)"
echo "800k SR=5: $(NPC_TILE_SR=5 timeout 200 $B --samples 800000 --variants 4992 2>&1 | tail -1)"
echo "== variant F: K = 1 up to 24 warps (72 registers)"
cp nimpress_b200/lib/libnimpress_cuda.so /tmp/keep.so; cp nimpress_b200/lib/variants/F.so nimpress_b200/lib/libnimpress_cuda.so
for n in 100000 500000 600000 700000 800000 900000; do
  v=$(( 4000000000 / n )); v=$(( v / 64 * 64 ))
  echo "F n=$n V=$v: $(timeout 200 $B --samples $n --variants $v 2>&1 | tail -1)"
done | tee gpurun_out/width_sweep3F.txt
cp /tmp/keep.so nimpress_b200/lib/libnimpress_cuda.so
echo "== int16"; timeout 300 python tools/bench_int16.py 2>&1 | tail -1 | cut -c1-400
