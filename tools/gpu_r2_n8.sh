#!/bin/bash
# round 2: 8 GPUs -- bench N=8 (NCCL combine inside the C ABI, parity check in the run) and the CLI over 8 devices from disk
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -12 > gpurun_out/topo_n8.txt; head -11 gpurun_out/topo_n8.txt | cut -c1-150
echo "== bench N=8"
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29688 bench.py --gpus 8 --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r2_n8.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_n8.json'))
print({k:d.get(k) for k in ('value','ms_per_step','parity_checked','n_gpus')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'fill', d['e2e']['with_host_fill']['ms_per_step'], d['roofline']['frac'], d['clocks'])
PY
echo "== CLI --devices=0-7, 2,000 variants x 500,000 samples from a BGZF BCF"
timeout 600 python tools/bench_cli_config3.py --variants 2000 --devices 0-7 --out gpurun_out/cli_config3_8dev.json > gpurun_out/cli3_8dev.log 2>&1; tail -c 1500 gpurun_out/cli3_8dev.log
echo "== multi-device tests on 8 GPUs"; timeout 600 python -m pytest tests/test_multi_device.py -x -q -m gpu 2>&1 | tail -3
