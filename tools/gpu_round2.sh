#!/bin/bash
mkdir -p gpurun_out
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== bench fused 16k"; timeout 300 python bench.py --variants 16384 --steps 3 --warmup 3 --cpu-seconds 5 2>&1 | tail -2 | tee gpurun_out/bench_fused_16k.json
echo "== bench two-kernel 16k"; NPC_FUSED=0 timeout 300 python bench.py --variants 16384 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -2 | tee gpurun_out/bench_2k_16k.json
for cfg in "4 7 2" "4 7 3" "2 12 3" "6 4 2" "3 9 3" "1 24 4"; do set -- $cfg
  echo "== fused R=$1 S=$2 L=$3"; NPC_FUSED_R=$1 NPC_FUSED_S=$2 NPC_FUSED_L=$3 timeout 200 python bench.py --variants 16384 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['config']['kernel_shape'])"
done 2>&1 | tee gpurun_out/sweep1.log
echo "== ncu full fused"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_fused' -s 2 -c 1 -o gpurun_out/prof_fused_r1 -f \
    python bench.py --variants 8192 --steps 1 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/ncu_fused.log 2>&1
ls -la gpurun_out | tail -5
