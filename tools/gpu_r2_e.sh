#!/bin/bash
# round 2, batch E: sanity after reverting the in-kernel combine, CLI from disk (config 2 and the config-3 slice), launch list
mkdir -p gpurun_out
echo "== parity"; timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_multi_device.py -x -q -m gpu 2>&1 | tail -3
echo "== shapes"; bash tools/gpu_shapes.sh 2>&1 | cut -c1-120
echo "== CLI config2"; timeout 600 python tools/bench_cli_config2.py --timing --out gpurun_out/cli_config2_r2.json 2>&1 | tail -40 | cut -c1-400
echo "== CLI config3 slice"; timeout 1500 python tools/bench_cli_config3.py --variants 10000 --out gpurun_out/cli_config3_r2.json > gpurun_out/cli3.log 2>&1; tail -c 2500 gpurun_out/cli3.log
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-extra > gpurun_out/ncu_launch_r2.log 2>&1; tail -1 gpurun_out/ncu_launch_r2.log | cut -c1-200
