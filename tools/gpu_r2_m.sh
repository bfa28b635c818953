#!/bin/bash
mkdir -p gpurun_out
echo "== multi-device + multi-score tests"; timeout 1500 python -m pytest tests/test_multi_device.py tests/test_multi_gpu.py tests/test_dosage_gpu.py -x -q -m gpu 2>&1 | tail -4
