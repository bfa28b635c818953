#!/bin/bash
mkdir -p gpurun_out
for s in 1 2; do timeout 400 python tools/fuzz_parity.py --cases 400 --seed $s --seconds 150 2>&1 | tail -2; done
NPC_TILE_A=1 timeout 300 python tools/fuzz_parity.py --cases 200 --seed 3 --seconds 90 2>&1 | tail -2
NPC_TILE_SC=10 NPC_TILE_SR=2 timeout 300 python tools/fuzz_parity.py --cases 200 --seed 4 --seconds 90 2>&1 | tail -2
NPC_TILE_K=2 timeout 300 python tools/fuzz_parity.py --cases 200 --seed 5 --seconds 90 2>&1 | tail -2
