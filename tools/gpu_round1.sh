#!/bin/bash
# First GPU visit: parity tests, a short bench, a launch list and one full ncu capture.
# Run as: gpurun --timeout 1500 -- 'bash tools/gpu_round1.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
free -g > gpurun_out/host.txt; nproc >> gpurun_out/host.txt
echo "== pytest -m gpu" ; timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "== bench small"; timeout 600 python bench.py --variants 16384 --steps 3 --warmup 3 --cpu-seconds 5 2>&1 | tail -3 | tee gpurun_out/bench_16k.json
echo "== ncu launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --variants 2048 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "== ncu full (accumulate + count)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_accum_i8x2|k_count_i8x2' -s 4 -c 2 -o gpurun_out/prof_r1 -f \
    python bench.py --variants 2048 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
