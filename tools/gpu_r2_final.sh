#!/bin/bash
# round 2: final validation of the tree
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== sanitizer (memcheck, long split everywhere)"; NPC_TILE_LONG_MB=1 bash tools/gpu_sanitize.sh 2>&1 | grep -v "^=========     and\|Race reported" | sed -n 1,16p
echo "== full bench"; timeout 1200 python bench.py 2>&1 | tail -1 > gpurun_out/bench_r2_final4.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_final4.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['roofline']['traffic'], d['clocks'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'fill', d['e2e']['with_host_fill']['ms_per_step'])
print({k:(v.get('roofline_frac'),v.get('launch_us'),v.get('call_ms')) for k,v in d['extra'].items()})
PY
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-300
