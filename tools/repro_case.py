import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import nimpress_b200 as nb
import orc
from util_cohort import assert_parity, random_cohort, random_rows
rng = np.random.default_rng(7)
n, V, n_rows = 157929, 85, 52
gt = random_cohort(rng, n, V, miss_rate=0.0, n_alt=9)
rows = random_rows(rng, V, n_rows=n_rows, n_alt=9)
for exact in (True, False):
    for block in (5, 52, 1, 3, 4, 8, 16, 17):
        eng = nb.Engine(n, max_rows_per_block=85, n_slots=2)
        eng.set_policy(imp_locus="ignore", imp_missing="ignore", imp_sample="fail", maxmis=0.02)
        eng.set_exact_order(exact); eng.reset()
        try:
            for r0 in range(0, n_rows, block):
                eng.score_host(gt, rows[r0:r0 + block])
            got = eng.finish()
            print("ok", exact, block, eng.kernel_shape)
        except Exception as e:
            print("FAIL", exact, block, eng.kernel_shape, repr(e)[:200])
        eng.close()
