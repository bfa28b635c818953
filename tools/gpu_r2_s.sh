#!/bin/bash
# round 2: the final validation of the tree + dosage fuzz
mkdir -p gpurun_out
echo "== dosage fuzz"; timeout 200 python tools/fuzz_parity.py --dosage --cases 400 --seed 21 --seconds 60 2>&1 | tail -2
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== full bench"; timeout 1200 python bench.py 2>&1 | tail -1 > gpurun_out/bench_r2_final.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_final.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['clocks'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'fill', d['e2e']['with_host_fill']['ms_per_step'])
print({k:(v.get('roofline_frac'),v.get('launch_us'),v.get('call_ms')) for k,v in d['extra'].items()})
PY
