#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu (parity only)" ; timeout 1800 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
B="python bench.py --variants 32768 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --brief"
echo "== default (tile4)"; timeout 200 $B 2>&1 | tail -1 | tee gpurun_out/sweep8.log
for cfg in "4 0 0 1" "4 0 0 2" "4 0 0 3" "4 0 0 4" "3 0 0 2" "5 0 0 2" "4 16 0 2" "4 0 12 2"; do set -- $cfg
  echo "== FAST SR=$1 SC=$2 L=$3 A=$4"; NPC_FAST_SR=$1 NPC_FAST_SC=$2 NPC_FAST_L=$3 NPC_FAST_A=$4 timeout 200 $B 2>&1 | tail -1
done 2>&1 | tee -a gpurun_out/sweep8.log
echo "== ncu full tile4"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_fused_tile4' -s 2 -c 1 -o gpurun_out/prof_tile4_r4 -f \
    python bench.py --variants 8192 --steps 1 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/ncu_fused.log 2>&1
tail -1 gpurun_out/ncu_fused.log
