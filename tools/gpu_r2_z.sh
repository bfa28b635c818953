#!/bin/bash
# round 2: 28-warp instance + a grid split of their own for long launches: parity, fuzz (threshold 1 MB: every launch "long"), sweep, bench
mkdir -p gpurun_out
echo "== parity"; timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -4
echo "== fuzz, long split everywhere"; NPC_TILE_LONG_MB=1 timeout 400 python tools/fuzz_parity.py --cases 500 --seed 36 --seconds 150 2>&1 | tail -2
echo "== fuzz"; timeout 300 python tools/fuzz_parity.py --cases 300 --seed 37 --seconds 80 2>&1 | tail -2
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --brief"
for n in 50000 100000 150000 200000 250000 300000 400000 500000 600000 660000 700000 740000 800000 900000 950000 1000000 1100000 1180000; do
  v=$(( 4000000000 / n )); v=$(( v / 64 * 64 ))
  echo "n=$n V=$v: $(timeout 200 $B --samples $n --variants $v 2>&1 | tail -1)"
done > gpurun_out/width_sweep_final2.txt; cut -c1-70 gpurun_out/width_sweep_final2.txt
echo "== full bench"; timeout 1200 python bench.py 2>&1 | tail -1 > gpurun_out/bench_r2_final3.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_final3.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['clocks'], d['config']['kernel_shape'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'fill', d['e2e']['with_host_fill']['ms_per_step'])
print({k:(v.get('roofline_frac'),v.get('launch_us'),v.get('call_ms')) for k,v in d['extra'].items()})
PY
echo "== bench, no long split"; NPC_TILE_LONG=0 timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-extra 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['clocks'])"
