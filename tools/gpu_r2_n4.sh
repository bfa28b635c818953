#!/bin/bash
mkdir -p gpurun_out
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29644 bench.py --gpus 4 --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r2_n4.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_n4.json'))
print({k:d.get(k) for k in ('value','ms_per_step','parity_checked','n_gpus')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['clocks'])
PY
echo "== reference arm under torchrun N=4"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29645 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-300
