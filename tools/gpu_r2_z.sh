#!/bin/bash
# round 2: ncu --set full of the bench launch shape after the long-launch split (74 slabs x 2 row groups, 27 warps), and the launch list
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused_pair -s 2 -c 1 -o gpurun_out/pair_long -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/ncu_long.log 2>&1; tail -2 gpurun_out/ncu_long.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_long.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/launches_long.log 2>&1; tail -1 gpurun_out/launches_long.log | cut -c1-200
