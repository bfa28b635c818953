#!/bin/bash
# compute-sanitizer over the multi-score contraction (tcgen05 / TMEM / TMA kernel + its preparation kernels).
mkdir -p gpurun_out
cat > /tmp/sanm.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, nimpress_b200 as nb, orc
from util_cohort import random_cohort, random_rows, assert_loci_equal
rng = np.random.default_rng(2)
for (n, V, S, pol) in [(5003, 150, 5, {}), (777, 70, 20, dict(imp_sample="fail"))]:
    gt = random_cohort(rng, n, V, miss_rate=0.03, n_alt=4, sentinel_rate=0.01)
    lists = [random_rows(rng, V, n_rows=int(rng.integers(1, 200)), n_alt=4) for _ in range(S)]
    offs = [float(k) for k in range(S)]
    eng = nb.Engine(n, max_rows_per_block=256, n_slots=2)
    eng.set_policy(**pol)
    assert eng.resident_reserve(V) >= V
    slot, view = eng.stage_acquire(); view[:V, :gt.shape[1]] = gt.view(np.uint8); eng.stage_upload(slot, V, 0)
    got = eng.score_resident_multi(lists, offs)
    assert eng.multi_contractions == 1
    for k in range(S):
        want = orc.score_matrix(gt, n, 2, lists[k], offset=offs[k], **pol)
        assert got[k][1] == want["nloci"]; assert_loci_equal(got[k][2], want["loci"])
        a, b = got[k][0], want["scores"]; ok = np.isfinite(b)
        assert np.array_equal(np.isnan(a), np.isnan(b)) and np.all(np.abs(a[ok] - b[ok]) <= 1e-12 * np.maximum(np.abs(b[ok]), 1e-3))
    eng.close()
    print("ok", n, V, S, pol, flush=True)
PY
for tool in memcheck racecheck synccheck; do
  echo "== $tool"; timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python /tmp/sanm.py > gpurun_out/sanitizer_multi_$tool.log 2>&1
  grep -E "^ok|ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" gpurun_out/sanitizer_multi_$tool.log | head -12
done
