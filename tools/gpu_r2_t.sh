#!/bin/bash
# round 2: sanitizer over the contraction after the converter changes + the CLI from disk with the new inflate-pool default
mkdir -p gpurun_out
bash tools/gpu_sanitize_multi.sh 2>&1 | tail -20
echo "== host tests"; timeout 600 python -m pytest tests/test_host_cpu.py tests/test_host_gpu.py -x -q 2>&1 | tail -2
echo "== CLI config3 slice"; timeout 900 python tools/bench_cli_config3.py --out gpurun_out/cli_config3_t.json 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['host_cores'], [ (round(r['wall_s'],3), round(r['gt_payload_gb_per_s'],2)) for r in d['runs']]); print(d['runs'][-1]['phases_ms'])"
