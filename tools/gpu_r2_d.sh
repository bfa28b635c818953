#!/bin/bash
# round 2, batch D: in-kernel partial-sum combine, short-launch trace, config-3 slice from disk through the CLI
mkdir -p gpurun_out
echo "== parity"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "== shapes"; bash tools/gpu_shapes.sh 2>&1 | grep -v "^2M\|two-kernel" 
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --brief"
echo "== config2 L=10"; NPC_TILE_L=10 timeout 200 $B --samples 100000 --variants 697 2>&1 | tail -1
echo "== config2 L=12 A=1"; NPC_TILE_L=12 NPC_TILE_A=1 timeout 200 $B --samples 100000 --variants 697 2>&1 | tail -1
echo "== config2 GR=4"; NPC_TILE_GR=4 timeout 200 $B --samples 100000 --variants 697 2>&1 | tail -1
echo "== config2 GR=3"; NPC_TILE_GR=3 timeout 200 $B --samples 100000 --variants 697 2>&1 | tail -1
echo "== trace"; timeout 300 python tools/trace_config2.py > gpurun_out/trace_r2_d.json 2>&1; python - <<'PY'
import json
for d in json.load(open('gpurun_out/trace_r2_d.json')):
    print(d['samples'], d['loci'], round(d['launch_us'],1), round(d['hbm_floor_us'],1), {k: round(v,1) for k,v in d['cta0_us_after_start'].items()})
PY
echo "== int16"; timeout 300 python tools/bench_int16.py 2>&1 | tail -1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_fused_pair -s 2 -c 1 -o gpurun_out/prof_r2_pair_int16 -f python tools/bench_int16.py --steps 1 > gpurun_out/ncu_int16.log 2>&1; tail -1 gpurun_out/ncu_int16.log | cut -c1-200
echo "== CLI config3 slice"; timeout 1500 python tools/bench_cli_config3.py --variants 10000 > gpurun_out/cli3.log 2>&1; tail -c 3000 gpurun_out/cli3.log
