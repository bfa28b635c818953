#!/bin/bash
# compute-sanitizer over a small parity run of every kernel path (tile4 with row groups, exact fused,
# two-kernel, generic widths).  Output: gpurun_out/sanitizer_*.log
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, nimpress_b200 as nb, orc
from util_cohort import random_cohort, random_rows, assert_parity
rng = np.random.default_rng(1)
for (n, V, width, ploidy, exact, env) in [(20011, 150, 1, 2, False, {}), (20011, 150, 1, 2, True, {}), (3001, 40, 2, 2, True, {}),
                                           (3001, 40, 1, 3, True, {}), (20011, 90, 1, 2, True, {"NPC_FUSED": "0"})]:
    for k, v in env.items(): os.environ[k] = v
    gt = random_cohort(rng, n, V, width=width, ploidy=ploidy, miss_rate=0.03, n_alt=5, sentinel_rate=0.01)
    rows = random_rows(rng, V, n_rows=V + 20, n_alt=5)
    eng = nb.Engine(n, ploidy=ploidy, gt_width=width, max_rows_per_block=256, n_slots=2)
    eng.set_policy(); eng.set_exact_order(exact); eng.reset()
    for r0 in range(0, len(rows), 97): eng.score_host(gt, rows[r0:r0 + 97])
    got = eng.finish(offset=0.5); shape = eng.kernel_shape; eng.close()
    assert_parity(got, orc.score_matrix(gt, n, ploidy, rows, offset=0.5), exact=exact or shape["fused"] != 2)
    print("ok", n, V, width, ploidy, exact, env, shape["fused"], flush=True)
    for k in env: os.environ.pop(k)
PY
for tool in memcheck racecheck synccheck; do
  echo "== $tool"; timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san.py > gpurun_out/sanitizer_$tool.log 2>&1
  grep -E "^ok|ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" gpurun_out/sanitizer_$tool.log | head -12
done
