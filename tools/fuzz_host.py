#!/usr/bin/env python3
"""Randomised end-to-end sweep of the host library on one GPU: trap-laden random datasets (tests/util_files.make_dataset:
duplicate sites, multi-base REF, multi-allelic records, FILTER values, haploid calls, rows out of order, loci absent / not
covered) as VCF text and BCF, random policies, with / without a BED, default and exact-order modes, one context or several
(NIMPRESS_SPLIT), forced rounds (NIMPRESS_SLAB_ROWS), with / without an index -- every result against the oracle's file run.

    python tools/fuzz_host.py [--seconds 150] [--seed 1]
"""
import argparse
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=150.0)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    import orc
    from nimpress_b200 import api
    from util_cohort import assert_loci_equal, bits, score_excess
    from util_files import make_dataset
    rng = np.random.default_rng(args.seed)
    t0 = time.time()
    done = 0
    while time.time() - t0 < args.seconds:
        tmp = tempfile.mkdtemp(prefix="fuzzhost_")
        n = int(np.exp(rng.uniform(np.log(2), np.log(3000))))
        V = int(rng.integers(6, 400))
        dt = [np.int8, np.int8, np.int16][rng.integers(3)]
        index = bool(rng.integers(0, 2))
        d = make_dataset(tmp, rng, n=n, V=V, miss_rate=float(rng.choice([0.0, 0.03, 0.2])), gt_dtype=dt, haploid_rate=float(rng.choice([0.0, 0.05])),
                         sorted_scores=bool(rng.integers(0, 2)), index=index, spread=int(rng.choice([5, 40, 400])))
        pol = dict(imp_locus=str(rng.choice(["ps", "homref", "fail", "ignore"])), imp_missing=str(rng.choice(["homref", "ignore"])),
                   imp_sample=str(rng.choice(["ps", "homref", "fail", "int_ps", "int_fail"])), maxmis=float(rng.choice([0.0, 0.05, 1.0])),
                   mincs=int(rng.choice([0, 100])), ignorefilt=bool(rng.integers(0, 2)))
        bed = d["bed"] if rng.integers(0, 2) else None
        exact = bool(rng.integers(0, 2))
        env = {}
        if rng.random() < 0.4:
            env["NIMPRESS_SPLIT"] = str(int(rng.integers(2, 6)))
        if rng.random() < 0.25:
            env["NIMPRESS_SLAB_ROWS"] = str(int(rng.integers(1, 40)))
        if index and rng.random() < 0.5:
            env["NIMPRESS_FORCE_INDEX"] = "1"
        geno = d["bcf"] if rng.integers(0, 2) else d["vcf"]
        desc = dict(n=n, V=V, dtype=np.dtype(dt).name, index=index, pol=pol, bed=bool(bed), exact=exact, env=env, geno=os.path.basename(geno))
        try:
            want = orc.compute_scores_files(d["score"], d["vcf"], bed, **pol)
            os.environ.update(env)
            got = api.run(d["score"], geno, bed, imp_locus=orc.LOCUS[pol["imp_locus"]], imp_missing=orc.MISSING[pol["imp_missing"]],
                          imp_sample=orc.SAMPLE[pol["imp_sample"]], maxmis=pol["maxmis"], mincs=pol["mincs"], ignorefilt=pol["ignorefilt"], exact_order=exact)
            assert got.samples == want["samples"] and got.nloci == want["nloci"], (got.nloci, want["nloci"])
            assert got.warnings == want["warn"]
            assert_loci_equal(got.loci, want["loci"])
            a, b = got.scores, want["scores"]
            assert np.array_equal(np.isnan(a), np.isnan(b))
            one_chain = exact and "NIMPRESS_SPLIT" not in env and got.rounds == 1
            if one_chain:
                ok = np.isfinite(b)
                assert np.array_equal(bits(a[ok]), bits(b[ok])), "exact-order scores differ in bits"
            else:
                assert score_excess(a, b, want) <= 1.0, score_excess(a, b, want)
        except Exception as e:               # noqa: BLE001
            print("FAIL", desc, repr(e)[:400], "files kept in", tmp, flush=True)
            raise SystemExit(1)
        finally:
            for k in env:
                os.environ.pop(k, None)
        shutil.rmtree(tmp, ignore_errors=True)
        done += 1
    print(f"host fuzz ok: {done} datasets in {time.time() - t0:.0f} s (seed {args.seed})")


if __name__ == "__main__":
    main()
