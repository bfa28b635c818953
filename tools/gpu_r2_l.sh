#!/bin/bash
mkdir -p gpurun_out
echo "== parity"; timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_host_gpu.py -x -q -m gpu 2>&1 | tail -3
echo "== int32 generic"; timeout 300 python tools/bench_int16.py --width 4 --variants 2048 2>&1 | tail -1 | cut -c1-330
echo "== --config 2"; timeout 300 python bench.py --config 2 2>&1 | tail -1 | cut -c1-500
