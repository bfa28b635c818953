#!/bin/bash
mkdir -p gpurun_out
echo "== multi tests"; timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -2
echo "== bench_multi"; NPC_TIMING=1 timeout 600 python tools/bench_multi.py --variants 20000 --reps 5 > gpurun_out/bench_multi_r2c.json 2> gpurun_out/bench_multi_r2c.err; cat gpurun_out/bench_multi_r2c.json | cut -c1-300; grep "contraction kernel" gpurun_out/bench_multi_r2c.err | tail -2
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_multi_contract' -s 2 -c 1 -o gpurun_out/prof_r2_multi_contract_b -f \
    python tools/bench_multi.py --variants 20000 --reps 2 > gpurun_out/ncu_multi_r2b.log 2>&1; tail -1 gpurun_out/ncu_multi_r2b.log | cut -c1-100
echo "== cfg2 probe"; timeout 300 python tools/cfg2_probe.py 2>&1 | tail -4
echo "== trace"; timeout 300 python tools/trace_config2.py > gpurun_out/trace_r2_j.json 2>&1; python - <<'PY'
import json
for d in json.load(open('gpurun_out/trace_r2_j.json')):
    print(d['samples'], d['loci'], round(d['launch_us'],1), round(d['hbm_floor_us'],1), {k: round(v,1) for k,v in d['cta0_us_after_start'].items()})
PY
