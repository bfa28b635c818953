#!/bin/bash
# round 2: final validation of the tree
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== sanitizer"; bash tools/gpu_sanitize.sh 2>&1 | grep -v "^=========     and\|Race reported" | tail -40
echo "== full bench"; timeout 1200 python bench.py 2>&1 | tail -1 > gpurun_out/bench_r2_final2.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_final2.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['clocks'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'fill', d['e2e']['with_host_fill']['ms_per_step'])
print({k:(v.get('roofline_frac'),v.get('launch_us'),v.get('call_ms')) for k,v in d['extra'].items()})
PY
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --brief"
for n in 50000 100000 150000 200000 250000 300000 400000 500000 600000 660000 700000 740000 800000 900000 1000000 1100000 1180000; do
  v=$(( 4000000000 / n )); v=$(( v / 64 * 64 ))
  echo "n=$n V=$v: $(timeout 200 $B --samples $n --variants $v 2>&1 | tail -1)"
done > gpurun_out/width_sweep_final.txt; cut -c1-70 gpurun_out/width_sweep_final.txt
