#!/bin/bash
mkdir -p gpurun_out
echo "== regression"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "shape_choice or wider or full_sample" 2>&1 | tail -2
for s in 1 11 12; do timeout 400 python tools/fuzz_parity.py --cases 500 --seed $s --seconds 120 2>&1 | tail -2; done
NPC_TILE_A=1 timeout 300 python tools/fuzz_parity.py --cases 300 --seed 13 --seconds 80 2>&1 | tail -2
echo "== shapes"; bash tools/gpu_shapes.sh 2>&1 | cut -c1-110
