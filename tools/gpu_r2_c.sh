#!/bin/bash
# round 2, batch C: FORMAT/DS rows, low-n launch changes, trace of short launches
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
B="python bench.py --variants 32768 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --brief"
echo "== v5"; timeout 200 $B 2>&1 | tail -1
echo "== shapes"; bash tools/gpu_shapes.sh
echo "== trace"; timeout 300 python tools/trace_config2.py > gpurun_out/trace_r2_c.json 2>&1; cat gpurun_out/trace_r2_c.json | head -120
