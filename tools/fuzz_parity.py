#!/usr/bin/env python3
"""Randomised parity sweep on one GPU: random cohort shapes (1 .. 300,000 samples, 1 .. 600 rows), storage widths, missing /
multi-allelic / sentinel rates, score-row mixes, policies, launch splits and kernel modes, each checked against the oracle
(records exactly; scores bit for bit in exact-order mode, at the re-association tolerance otherwise).

    python tools/fuzz_parity.py [--cases 300] [--seed 1]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=300)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--seconds", type=float, default=240.0)
    ap.add_argument("--dosage", action="store_true", help="FORMAT/DS rows (fp32 dosages) instead of GT rows")
    args = ap.parse_args()
    import nimpress_b200 as nb
    import orc
    from util_cohort import assert_parity, random_cohort, random_rows
    rng = np.random.default_rng(args.seed)
    t0 = time.time()
    done = 0
    for case in range(args.cases):
        if time.time() - t0 > args.seconds:
            break
        n = int(np.exp(rng.uniform(0, np.log(300_000))))
        V = int(np.exp(rng.uniform(0, np.log(600))))
        width = int(rng.choice([1, 1, 1, 2]))
        n_alt = int(rng.choice([1, 1, 2, 3, 9]))
        miss = float(rng.choice([0.0, 0.005, 0.05, 0.3]))
        sent = float(rng.choice([0.0, 0.0, 0.01]))
        exact = bool(rng.integers(0, 2))
        pol = dict(imp_locus=str(rng.choice(["ps", "homref", "fail", "ignore"])), imp_missing=str(rng.choice(["homref", "ignore"])),
                   imp_sample=str(rng.choice(["ps", "homref", "fail", "int_ps", "int_fail"])), maxmis=float(rng.choice([0.0, 0.02, 0.05, 1.0])),
                   mincs=int(rng.choice([0, 100, n + 1])))
        n_rows = int(V * rng.uniform(0.5, 2.0)) + 1
        if args.dosage:
            n_alt, width = 1, 4
            gt = np.zeros((V, -(-n // 32) * 32), dtype=np.float32)
            gt[:, :n] = np.round(rng.uniform(0, 2, size=(V, n)) * rng.choice([0.0, 1.0, 1.0], size=(V, 1)), int(rng.integers(1, 7)))
            b = gt.view(np.uint32)
            m = rng.random((V, n)) < miss
            b[:, :n][m] = rng.choice(np.array([0x7F800001, 0x7F800002, 0x7FC00000], dtype=np.uint32), size=int(m.sum()))
            rows = random_rows(rng, V, n_rows=n_rows)
            rows["eaidx"] = np.where(rows["ref_is_ea"] == 1, 0, 1)
        else:
            gt = random_cohort(rng, n, V, width=width, miss_rate=miss, n_alt=n_alt, sentinel_rate=sent)
            rows = random_rows(rng, V, n_rows=n_rows, n_alt=n_alt)
        block = int(rng.choice([n_rows, max(1, n_rows // 3), 5]))
        offset = float(rng.choice([0.0, 0.5, -3.25]))
        desc = dict(case=case, n=n, V=V, width=width, n_alt=n_alt, miss=miss, sent=sent, exact=exact, pol=pol, n_rows=n_rows, block=block)
        try:
            eng = nb.Engine(n, ploidy=1 if args.dosage else 2, gt_width=width, max_rows_per_block=max(n_rows, V, 1), n_slots=2)
            if args.dosage:
                eng.set_dosage_rows(True)
            eng.set_policy(**pol); eng.set_exact_order(exact); eng.reset()
            for r0 in range(0, n_rows, block):
                eng.score_host(gt, rows[r0:r0 + block])
            got = eng.finish(offset=offset)
            shape = eng.kernel_shape
            eng.close()
            want = orc.score_matrix(gt, n, 1 if args.dosage else 2, rows.astype(orc.ROW_DTYPE), offset=offset, **pol)
            assert_parity(got, want, exact=exact or shape["fused"] == 0 or args.dosage)
        except Exception as e:               # noqa: BLE001
            print("FAIL", desc, repr(e)[:300], flush=True)
            raise SystemExit(1)
        done += 1
    print(f"fuzz ok: {done} cases in {time.time() - t0:.0f} s (seed {args.seed})")


if __name__ == "__main__":
    main()
