#!/bin/bash
# round 2, final tree: the 8-GPU bench line (variant-sharded, NCCL combine inside the C ABI, parity checked in the run)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29688 bench.py --gpus 8 --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r2_n8_final.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_n8_final.json'))
print({k:d.get(k) for k in ('value','ms_per_step','parity_checked','n_gpus','gpu_launches')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['clocks'])
PY
