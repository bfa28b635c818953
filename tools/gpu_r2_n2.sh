#!/bin/bash
# round 2: the multi-GPU product path on 2 GPUs -- npc_reduce, --devices, npc_comm_* (NCCL), bench N=2 with the parity check
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -8
echo "== pytest tests/test_multi_device.py"; timeout 1200 python -m pytest tests/test_multi_device.py -x -q -m gpu 2>&1 | tail -8
echo "== bench N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r2_n2.json
cut -c1-400 gpurun_out/bench_r2_n2.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_n2.json'))
print({k:d.get(k) for k in ('value','ms_per_step','parity_checked','n_gpus')}, d['e2e']['value'], d['e2e']['ms_per_step'])
PY
