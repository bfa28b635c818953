#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --variants 16384 --steps 5 --warmup 3 --no-cpu-baseline --no-extra"
echo "== default pinned"; timeout 300 $B 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['with_host_fill']['ms_per_step'])"
echo "== write-combined"; NPC_STAGE_WC=1 timeout 300 $B 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['with_host_fill']['ms_per_step'])"
