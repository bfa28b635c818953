#!/bin/bash
mkdir -p gpurun_out
echo "== multi tests"; timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_multi_device.py -x -q -m gpu 2>&1 | tail -2
echo "== bench_multi"; NPC_TIMING=1 timeout 600 python tools/bench_multi.py --reps 5 > gpurun_out/bench_multi_r2d.json 2> gpurun_out/bench_multi_r2d.err; cat gpurun_out/bench_multi_r2d.json | cut -c1-300; grep "npc multi" gpurun_out/bench_multi_r2d.err | tail -7
