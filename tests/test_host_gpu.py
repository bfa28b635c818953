"""GPU tests of the whole product path: C++ host (readers, streaming findVariant, staging into
the resident slab) -> libnimpress_cuda -> scores, against the oracle run on the same files."""
import os
import subprocess

import numpy as np
import pytest

import orc
from util_cohort import assert_loci_equal, bits, score_excess
from util_files import make_dataset

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
S1 = os.path.join(ROOT, "tests", "golden", "set1")
LOC = ("ps", "homref", "fail", "ignore"); MIS = ("homref", "ignore"); SAM = ("ps", "homref", "fail", "int_ps", "int_fail")


@pytest.fixture(scope="module")
def api():
    import __graft_entry__ as g
    g.build()
    from nimpress_b200 import api
    return api


def compare(api, score, geno, vcf_for_oracle, bed, exact, **pol):
    want = orc.compute_scores_files(score, vcf_for_oracle, bed, **pol)
    got = api.run(score, geno, bed, imp_locus=orc.LOCUS[pol.get("imp_locus", "ps")], imp_missing=orc.MISSING[pol.get("imp_missing", "homref")],
                  imp_sample=orc.SAMPLE[pol.get("imp_sample", "int_ps")], maxmis=pol.get("maxmis", 0.05), afmisp=pol.get("afmisp", 0.001),
                  mincs=pol.get("mincs", 100), ignorefilt=pol.get("ignorefilt", False), exact_order=exact)
    assert got.samples == want["samples"] and got.nloci == want["nloci"] and got.rounds == 1
    assert_loci_equal(got.loci, want["loci"])
    a, b = got.scores, want["scores"]
    assert np.array_equal(np.isnan(a), np.isnan(b))
    ok = np.isfinite(b)
    if exact:
        assert np.array_equal(bits(a[ok]), bits(b[ok]))
    else:
        assert score_excess(a, b, want) <= 1.0
    assert got.warnings == want["warn"]
    return got


@pytest.mark.parametrize("exact", [False, True], ids=["tile4", "exact-order"])
def test_set1_all_policies_vs_oracle(api, exact):
    """Every policy combination on the reference fixture: scores, per-locus records and the WARN
    text (incl. the AF-mismatch binomial test at the default --afmisp) equal the oracle's."""
    sc, vc, bed = (os.path.join(S1, f) for f in ("set1.score", "set1.vcf.gz", "set1.bed"))
    for l in LOC:
        for m in MIS:
            for s in SAM:
                for mm, cs in ((1.0, 100), (0.2, 3), (0.05, 0)):
                    for b in (None, bed):
                        compare(api, sc, vc, vc, b, exact, imp_locus=l, imp_missing=m, imp_sample=s, maxmis=mm, mincs=cs)
    compare(api, sc, vc, vc, None, exact, ignorefilt=True, afmisp=1.0)


@pytest.mark.parametrize("seed,exact", [(11, True), (12, False), (13, True)])
def test_synthetic_files_vcf_and_bcf(api, tmp_path, seed, exact):
    """Synthetic cohort with the lookup corner cases, as BGZF VCF text and as BCF: both equal the
    oracle (which reads the VCF text), under several policies, with and without --cov."""
    rng = np.random.default_rng(seed)
    d = make_dataset(str(tmp_path), rng, n=1203, V=300, sorted_scores=seed != 13)
    for geno in (d["vcf"], d["bcf"]):
        for bed in (None, d["bed"]):
            compare(api, d["score"], geno, d["vcf"], bed, exact)
            compare(api, d["score"], geno, d["vcf"], bed, exact, imp_locus="homref", imp_sample="fail", maxmis=0.02, afmisp=0.05)
            compare(api, d["score"], geno, d["vcf"], bed, exact, imp_locus="ignore", imp_missing="ignore", imp_sample="int_fail", mincs=1100,
                    ignorefilt=True)


def test_wider_layouts_restart(api, tmp_path):
    """A matched record with int16 GT or ploidy 3 makes the host rerun the pass with that layout
    (generic kernels): results still equal the oracle bit for bit."""
    from util_bcf import write_bcf, write_vcf
    rng = np.random.default_rng(21)
    d = make_dataset(str(tmp_path), rng, n=211, V=60, gt_dtype=np.int16)
    compare(api, d["score"], d["bcf"], d["vcf"], None, True)
    d = make_dataset(str(tmp_path), rng, n=97, V=45, ploidy=3, haploid_rate=0.1)
    compare(api, d["score"], d["bcf"], d["vcf"], d["bed"], True)
    compare(api, d["score"], d["vcf"], d["vcf"], None, True)


def test_cli_stdout_matches_oracle_cli(api, tmp_path):
    """`nimpress [options] <scoredef> <genotypes>`: stdout (WARN lines then sample<TAB>score in the
    reference's float format) and exit codes, against the oracle's CLI."""
    exe = os.path.join(ROOT, "nimpress_b200", "bin", "nimpress")
    orc_exe = os.path.join(ROOT, "oracle", "_build", "nimpress_oracle")
    sc, vc, bed = (os.path.join(S1, f) for f in ("set1.score", "set1.vcf.gz", "set1.bed"))
    for opts in ([], ["--imp-sample=fail", "--maxmis=0.2"], [f"--cov={bed}", "--imp-locus=homref", "--mincs=3"],
                 ["--ignorefilt", "--imp-locus=ignore", "--imp-missing=ignore", "--maxmis=1.0", "--mincs=0", "--afmisp=1.0"]):
        a = subprocess.run([exe, "--exact-order", *opts, sc, vc], capture_output=True, text=True)
        b = subprocess.run([orc_exe, *opts, sc, vc], capture_output=True, text=True)
        assert a.returncode == 0 and a.stdout == b.stdout, (opts, a.stdout, b.stdout, a.stderr)
    rng = np.random.default_rng(3)
    d = make_dataset(str(tmp_path), rng, n=400, V=150)
    a = subprocess.run([exe, "--exact-order", f"--cov={d['bed']}", d["score"], d["bcf"]], capture_output=True, text=True)
    b = subprocess.run([orc_exe, f"--cov={d['bed']}", d["score"], d["vcf"]], capture_output=True, text=True)
    assert a.returncode == 0 and a.stdout == b.stdout
    assert subprocess.run([exe, sc, "/no/such.vcf"], capture_output=True, text=True).returncode == 255
    assert subprocess.run([exe, "/no/such.score", vc], capture_output=True, text=True).returncode == 255
    assert subprocess.run([exe, "--imp-locus=bogus", sc, vc], capture_output=True, text=True).returncode == 1
    v = subprocess.run([exe, "--version"], capture_output=True, text=True)
    assert v.stdout.strip() == "nimpress 1.0.0"


def test_rounds_when_slab_is_too_small(api, tmp_path, monkeypatch):
    """Matched genotype rows exceeding the resident slab are scored in rounds (NIMPRESS_SLAB_ROWS forces
    it): per-locus records still bit-equal, scores within 1e-12 (terms are added in a different order),
    WARN text unchanged; nph_result_rounds reports the rounds."""
    rng = np.random.default_rng(31)
    d = make_dataset(str(tmp_path), rng, n=1500, V=300, sorted_scores=False)
    want = orc.compute_scores_files(d["score"], d["vcf"], d["bed"])
    for cap in ("7", "64"):
        monkeypatch.setenv("NIMPRESS_SLAB_ROWS", cap)
        got = api.run(d["score"], d["bcf"], d["bed"], exact_order=True)
        assert got.rounds > 2 and got.nloci == want["nloci"] and got.warnings == want["warn"]
        assert_loci_equal(got.loci, want["loci"])
        a, b = got.scores, want["scores"]
        assert np.array_equal(np.isnan(a), np.isnan(b))
        ok = np.isfinite(b)
        assert score_excess(a, b, want) <= 1.0
    monkeypatch.delenv("NIMPRESS_SLAB_ROWS")
    got = api.run(d["score"], d["bcf"], d["bed"], exact_order=True)
    assert got.rounds == 1 and np.array_equal(bits(got.scores[np.isfinite(got.scores)]), bits(want["scores"][np.isfinite(want["scores"])]))


def test_indexed_files_are_read_by_region(api, tmp_path, monkeypatch):
    """A sparse score over files with a .tbi / .csi next to them: the pass jumps from locus to locus (index_seeks > 0,
    a fraction of the records read) and scores, per-locus records and WARN text equal the oracle's on the whole file;
    several score files at once share the one indexed pass."""
    import util_bcf
    rng = np.random.default_rng(41)
    monkeypatch.setattr(util_bcf, "INDEX_BLOCK", 6000)
    d = make_dataset(str(tmp_path), rng, n=400, V=900, index=True, spread=300)
    lines = open(d["score"]).read().split("\n")
    head, ents = lines[:5], [e for e in lines[5:] if e]
    sparse = []
    for k in range(3):
        keep = [e for i, e in enumerate(ents) if rng.random() < 0.04 or i >= len(ents) - 10]
        p = tmp_path / f"sparse{k}.score"
        p.write_text("\n".join(head + keep) + "\n")
        sparse.append(str(p))
    monkeypatch.setenv("NIMPRESS_FORCE_INDEX", "1")
    for f in (d["vcf"], d["bcf"]):
        want = orc.compute_scores_files(sparse[0], d["vcf"], d["bed"])
        got = api.run(sparse[0], f, d["bed"], exact_order=True)
        assert got.index_seeks > 0 and got.records_read < len(d["records"]) // 2
        assert got.nloci == want["nloci"] and got.warnings == want["warn"]
        assert_loci_equal(got.loci, want["loci"])
        ok = np.isfinite(want["scores"])
        assert np.array_equal(bits(got.scores[ok]), bits(want["scores"][ok]))
    multi = api.run_multi(sparse, d["bcf"], exact_order=True)
    for k, p in enumerate(sparse):
        want = orc.compute_scores_files(p, d["vcf"])
        assert multi[k].index_seeks > 0 and multi[k].warnings == want["warn"]
        assert_loci_equal(multi[k].loci, want["loci"])
        ok = np.isfinite(want["scores"])
        assert np.array_equal(bits(multi[k].scores[ok]), bits(want["scores"][ok]))
    monkeypatch.setenv("NIMPRESS_NO_INDEX", "1")
    monkeypatch.delenv("NIMPRESS_FORCE_INDEX")
    assert api.run(sparse[0], d["bcf"]).index_seeks == 0
