"""The C oracle against an independent pure-Python restatement written in the reference's own shape
(tests/pyref.py), on random small cohorts over every policy combination, storage width and ploidy:
scores bit for bit, per-locus records field by field.  CPU only.  Together with the golden vectors
(test_oracle_golden.py) this is what the parity claims of the GPU tests rest on."""
import math
import struct

import numpy as np
import pytest

import orc
import pyref
from util_cohort import random_cohort, random_rows

LOC = ("ps", "homref", "fail", "ignore"); MIS = ("homref", "ignore"); SAM = ("ps", "homref", "fail", "int_ps", "int_fail")


def same_bits(a, b):
    return struct.pack("<d", a) == struct.pack("<d", b) or (math.isnan(a) and math.isnan(b))


def compare(gt, n, ploidy, rows, offset, **pol):
    want_s, want_n, want_l = pyref.score(gt, n, ploidy, rows, offset, **pol)
    got = orc.score_matrix(gt, n, ploidy, rows, offset=offset, **pol)
    assert got["nloci"] == want_n
    for i, (a, b) in enumerate(zip(got["scores"], want_s)):
        assert same_bits(float(a), b), (i, a, b, pol)
    for r, (L, w) in enumerate(zip(got["loci"], want_l)):
        klass, used, ngt, nmiss, neff, imp = w
        assert (L["klass"], L["used"]) == (klass, used), (r, L, w)
        if ngt >= 0:
            assert (L["ngt"], L["nmiss"], L["neff"]) == (ngt, nmiss, neff), (r, L, w)
        if used:
            assert same_bits(float(L["imputed"]), imp), (r, L, w)


@pytest.mark.parametrize("seed", range(6))
def test_all_policies_small_cohorts(seed):
    rng = np.random.default_rng(1000 + seed)
    n, V = int(rng.integers(1, 40)), int(rng.integers(1, 12))
    gt = random_cohort(rng, n, V, miss_rate=float(rng.uniform(0, 0.3)), n_alt=3, sentinel_rate=0.05, invalid_rate=0.02)
    rows = random_rows(rng, V, n_rows=V + 5, n_alt=3, nan_eaf_rate=0.1)
    for l in LOC:
        for m in MIS:
            for s in SAM:
                for maxmis, mincs in ((0.05, 100), (0.2, 3), (1.0, 0), (0.0, n)):
                    compare(gt, n, 2, rows, 0.125, imp_locus=l, imp_missing=m, imp_sample=s, maxmis=maxmis, mincs=mincs)


@pytest.mark.parametrize("width,ploidy", [(1, 1), (1, 3), (2, 2), (4, 2), (2, 4)])
def test_widths_and_ploidies(width, ploidy):
    rng = np.random.default_rng(width * 10 + ploidy)
    n, V = 23, 9
    gt = random_cohort(rng, n, V, width=width, ploidy=ploidy, miss_rate=0.1, n_alt=4, sentinel_rate=0.1)
    rows = random_rows(rng, V, n_rows=14, n_alt=4)
    for s in SAM:
        compare(gt, n, ploidy, rows, -1.5, imp_sample=s, maxmis=0.3, mincs=5)


def test_no_locus_used_gives_nan():
    """nloci == 0: 0/0 (:643-649)."""
    rng = np.random.default_rng(3)
    gt = random_cohort(rng, 7, 2)
    rows = random_rows(rng, 2, n_rows=4, kinds=(0, 0.5, 0.5, 0))
    compare(gt, 7, 2, rows, 0.5, imp_locus="ignore", imp_missing="ignore")
