#!/bin/bash
# round 2, batch F: multi-score call after the host-side overlap, start-up cost of a CUDA process on this box, all GPU tests
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "== bench_multi"; NPC_TIMING=1 timeout 600 python tools/bench_multi.py --reps 5 > gpurun_out/bench_multi_r2.json 2> gpurun_out/bench_multi_r2.err; cat gpurun_out/bench_multi_r2.json | cut -c1-700; grep "npc multi" gpurun_out/bench_multi_r2.err | tail -12
echo "== process start-up"
for i in 1 2 3; do python - <<'PY'
import ctypes, time
t0 = time.perf_counter(); L = ctypes.CDLL("nimpress_b200/lib/libnimpress_cuda.so"); t1 = time.perf_counter()
L.npc_warmup(0); t2 = time.perf_counter()
print(f"dlopen libnimpress_cuda.so {1e3*(t1-t0):.0f} ms, npc_warmup(0) = cudaSetDevice + cudaFree(0): {1e3*(t2-t1):.0f} ms")
PY
done
nvidia-smi --query-gpu=persistence_mode --format=csv
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
