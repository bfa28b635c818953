#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
B="python bench.py --variants 32768 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --brief"
echo "== default"; timeout 200 $B 2>&1 | tail -1 | tee gpurun_out/sweep4.log
for cfg in "4 3 0 0 4" "4 2 0 0 4" "4 4 0 0 4" "2 5 0 0 4" "8 2 0 0 4" "8 1 0 0 4" "4 3 0 0 2" "4 3 0 0 6" "3 3 0 0 4" "4 3 0 4 4"; do set -- $cfg
  echo "== R=$1 SR=$2 SC=$3 L=$4 A=$5"; NPC_FUSED_R=$1 NPC_FUSED_SR=$2 NPC_FUSED_SC=$3 NPC_FUSED_L=$4 NPC_FUSED_A=$5 timeout 200 $B 2>&1 | tail -1
done 2>&1 | tee -a gpurun_out/sweep4.log
echo "== ncu full fused"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_fused' -s 2 -c 1 -o gpurun_out/prof_fused_r4 -f \
    python bench.py --variants 8192 --steps 1 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/ncu_fused.log 2>&1
tail -2 gpurun_out/ncu_fused.log
