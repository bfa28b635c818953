"""Pins the CPU oracle against every golden vector the reference's own tests hold for the
scoring path: the 13 cases of tests/test_set1.nim (incl. the PLINK 1.90 golden) and the
known answers of tests/test_stats.nim.  CPU only."""
import json
import math
import os

import numpy as np
import pytest

import orc

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
S1 = os.path.join(G, "set1")
CASES = json.load(open(os.path.join(G, "set1_expected.json")))["cases"]
KATS = json.load(open(os.path.join(G, "stats_kat.json")))["kats"]


def expected_vector(case):
    out = []
    for e in case["expected"]:
        if e is None:
            out.append(math.nan)
        elif isinstance(e, list):
            out.append(e[1] - e[2])
        else:
            out.append(e)
    return np.array(out)


def run_case(case):
    return orc.compute_scores_files(
        os.path.join(S1, "set1.score"), os.path.join(S1, "set1.vcf.gz"),
        os.path.join(S1, "set1.bed") if case["cov"] else None,
        imp_locus=case["imp_locus"], imp_missing=case["imp_missing"], imp_sample=case["imp_sample"],
        maxmis=case["maxmis"], afmisp=case["afmisp"], mincs=case["mincs"], ignorefilt=case["ignorefilt"])


@pytest.mark.parametrize("case", CASES, ids=[c["ref"] for c in CASES])
def test_set1_vectors(case):
    """checkFloats of tests/test_set1.nim:14-22: NaN pattern equal, |x - t| <= 1e-4."""
    got = run_case(case)["scores"]
    exp = expected_vector(case)
    assert len(got) == 6
    assert np.array_equal(np.isnan(got), np.isnan(exp))
    ok = ~np.isnan(exp)
    assert np.all(np.abs(got[ok] - exp[ok]) <= 1e-4)


def test_plink190_golden():
    """tests/set1.plink190.result:2-7 -- SCORE column + 0.123 offset (tests/test_set1.nim:180-190)."""
    plink = [float(l.split()[5]) for l in open(os.path.join(S1, "set1.plink190.result")).read().splitlines()[1:]]
    case = [c for c in CASES if c["imp_locus"] == "ignore"][0]
    got = run_case(case)["scores"]
    assert np.all(np.abs(got - (0.123 + np.array(plink))) <= 1e-4)


def test_set1_locus_records():
    """Per-locus facts of SURVEY.md Appendix B (derived from the fixture by hand)."""
    r = run_case(CASES[0])  # ps, homref, fail, maxmis 1.0, no cov
    loci = r["loci"]
    assert list(loci["klass"]) == [orc.CLASS_OK, orc.CLASS_FILTER, orc.CLASS_ABSENT, orc.CLASS_OK, orc.CLASS_OK,
                                   orc.CLASS_OK]
    assert list(loci["eaidx"]) == [0, 1, -1, 2, 1, 1]
    assert list(loci["ngt"]) == [5, -1, -1, 5, 1, 5]
    assert list(loci["nmiss"]) == [1, -1, -1, 1, 5, 1]
    assert list(loci["neff"]) == [7, -1, -1, 2, 0, 7]
    assert r["nloci"] == 6 and r["samples"] == ["S1", "S2", "S3", "S4", "S5", "S6"]
    r = run_case(CASES[1])  # maxmis 0.2
    assert list(r["loci"]["klass"]) == [orc.CLASS_OK, orc.CLASS_FILTER, orc.CLASS_ABSENT, orc.CLASS_OK,
                                        orc.CLASS_MAXMIS, orc.CLASS_OK]
    r = run_case([c for c in CASES if c["cov"]][0])
    assert list(r["loci"]["klass"]) == [orc.CLASS_NOTCOV, orc.CLASS_FILTER, orc.CLASS_ABSENT, orc.CLASS_NOTCOV,
                                        orc.CLASS_NOTCOV, orc.CLASS_OK]


def test_set1_full_precision_appendix_b():
    """Last-digit values of SURVEY.md Appendix B: summation order == reference order."""
    expect = {
        1: "0.07516666666666666 0.1085 nan nan nan -0.01649999999999999",
        3: "0.07516666666666666 0.108 0.07016666666666665 0.03683333333333333 0.006833333333333316 -0.01649999999999999",
        11: "0.08099999999999999 0.08099999999999999 0.08099999999999999 0.1545 0.006000000000000005 0.006000000000000005",
        12: "0.093 0.113 0.047 0.027 -0.009000000000000008 -0.03700000000000001",
    }
    for i, txt in expect.items():
        got = " ".join(orc.format_float(x) for x in run_case(CASES[i])["scores"])
        assert got == txt


@pytest.mark.parametrize("kat", KATS, ids=[f'{k["fn"]}@{k["ref"].split(":")[-1]}' for k in KATS])
def test_stats_kat(kat):
    """check_floatvalue of tests/test_stats.nim:6-17: rel 1e-5, abs 1e-9 for tiny targets."""
    L = orc.lib()
    fn = {"dbinom": L.orc_dbinom, "pbinom": L.orc_pbinom, "binom_test": L.orc_binom_test, "betai": L.orc_betai}[kat["fn"]]
    a = kat["args"]
    val = fn(a[0], a[1], a[2]) if kat["fn"] == "betai" else fn(int(a[0]), int(a[1]), a[2])
    t = kat["target"]
    if kat["kind"] == "exact":
        assert val == t
    elif abs(t) < 1e-9:
        assert abs(val - t) < 1e-9
    else:
        assert abs((val - t) / t) < 1e-5


def test_float_format_corpus():
    """The 3,528 scores the reference printed (scores/*_nimpress_res.txt) survive
    parse -> orc_format_float unchanged: the `$`(float) == "%.16g" (+".0") restatement."""
    vals = open(os.path.join(G, "res_format_corpus.tsv")).read().split()
    assert len(vals) == 3528
    bad = [v for v in vals if orc.format_float(float(v)) != v]
    assert not bad, bad[:5]
    assert orc.format_float(2.0) == "2.0" and orc.format_float(float("nan")) == "nan"
    assert orc.format_float(1e22) == "1e+22" and orc.format_float(-math.inf) == "-inf"
