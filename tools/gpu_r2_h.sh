#!/bin/bash
# round 2, batch H: all GPU tests (wide cohorts, more contexts than rows), wide-cohort rate, the final default bench line
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --brief"
echo "== 2M samples x 2048 (decided-mode slabs)"; timeout 300 $B --samples 2000000 --variants 2048 2>&1 | tail -1
echo "== 2M samples x 2048, NPC_WIDE=0 (two-kernel)"; NPC_WIDE=0 timeout 300 $B --samples 2000000 --variants 2048 2>&1 | tail -1
echo "== 2M exact"; NPC_EXACT=1 timeout 300 $B --samples 2000000 --variants 2048 2>&1 | tail -1
echo "== 1.5M samples x 4096"; timeout 300 $B --samples 1500000 --variants 4096 2>&1 | tail -1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== full bench"; timeout 1200 python bench.py 2>&1 | tail -1 > gpurun_out/bench_r2_final.json; cut -c1-300 gpurun_out/bench_r2_final.json
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_r2_reference.json; cut -c1-300 gpurun_out/bench_r2_reference.json
