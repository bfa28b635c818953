"""GPU tests of the multi-score path (BASELINE.json configs[3]): several score files over ONE pass of
the genotype file and one resident slab.  The reference has no such mode -- it is one nimpress run
per score file -- so the expectation for file k is the oracle run on file k alone."""
import os
import subprocess

import numpy as np
import pytest

import orc
from util_cohort import assert_loci_equal, bits, random_cohort, random_rows
from util_files import derive_scores, make_dataset, write_score
from util_bcf import write_bcf, write_vcf
from util_vcf import read_score

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
REAL = sorted(os.path.join(G, "scores", f) for f in os.listdir(os.path.join(G, "scores")) if f.endswith(".scores"))


@pytest.fixture(scope="module")
def api():
    import __graft_entry__ as g
    g.build()
    from nimpress_b200 import api
    return api


@pytest.fixture(scope="module")
def nb():
    import __graft_entry__ as g
    g.build()
    import nimpress_b200 as nb
    return nb


def check(got, want, exact, rtol=1e-12):
    assert got.samples == want["samples"] and got.nloci == want["nloci"]
    assert_loci_equal(got.loci, want["loci"])
    a, b = got.scores, want["scores"]
    assert np.array_equal(np.isnan(a), np.isnan(b))
    ok = np.isfinite(b)
    if exact:
        assert np.array_equal(bits(a[ok]), bits(b[ok]))
    else:
        assert np.all(np.abs(a[ok] - b[ok]) <= rtol * np.maximum(np.abs(b[ok]), 1e-3))
    assert got.warnings == want["warn"]


@pytest.mark.parametrize("exact", [False, True], ids=["default", "exact-order"])
def test_multi_equals_one_run_per_file(api, tmp_path, exact):
    """Seven score files (the trap-laden original + six derived: other effect alleles, repeated and
    shuffled rows) in one pass == the oracle on each file alone, under several policies."""
    rng = np.random.default_rng(5)
    d = make_dataset(str(tmp_path), rng, n=700, V=240, sorted_scores=False)
    paths = [d["score"]] + derive_scores(str(tmp_path), rng, d["entries"], 6)
    for pol in (dict(), dict(imp_locus="fail", imp_sample="int_fail", maxmis=0.02, mincs=800),
                dict(imp_locus="ignore", imp_missing="ignore", imp_sample="homref", ignorefilt=True)):
        for bed in (None, d["bed"]):
            got = api.run_multi(paths, d["bcf"], bed, imp_locus=orc.LOCUS[pol.get("imp_locus", "ps")],
                                imp_missing=orc.MISSING[pol.get("imp_missing", "homref")],
                                imp_sample=orc.SAMPLE[pol.get("imp_sample", "int_ps")], maxmis=pol.get("maxmis", 0.05),
                                mincs=pol.get("mincs", 100), ignorefilt=pol.get("ignorefilt", False), exact_order=exact)
            assert len(got) == len(paths)
            for k, p in enumerate(paths):
                assert got[k].rounds == 1
                check(got[k], orc.compute_scores_files(p, d["vcf"], bed, **pol), exact)


def config4_files(tmp, rng, n, n_synth=14):
    """BASELINE.json configs[3] as SURVEY.md section 8(d) makes it concrete: the 4 bundled score
    definitions + 14 synthetic ones over the union of their sites, one synthetic cohort."""
    sites = {}
    for p in REAL:
        for e in read_score(p)[1]:
            sites.setdefault((e["contig"], e["pos"], e["ref"]), set()).add(e["ea"])
    order = {str(c): i for i, c in enumerate(list(range(1, 23)) + ["X", "Y"])}
    keys = sorted(sites, key=lambda k: (order.get(k[0], 99), k[1], k[2]))
    samples = [f"S{i + 1}" for i in range(n)]
    records, entries = [], []
    for (c, pos, ref) in keys:
        alts = sorted(a for a in sites[(c, pos, ref)] if a != ref) or ["N"]
        af = rng.uniform(0.02, 0.5)
        al = (rng.random((n, 2)) < af) * rng.integers(1, len(alts) + 1, size=(n, 2))
        g = ((al + 1) << 1).astype(np.int8)
        g[rng.random(n) < 0.005] = 0
        records.append(dict(contig=c, pos=pos, ref=ref, alts=alts, filter="PASS", gt=g))
        entries.append((c, pos, ref, alts[0], 0.0, round(float(af), 4)))
    contigs = sorted({k[0] for k in keys}, key=lambda c: order.get(c, 99))
    vcf = write_vcf(os.path.join(tmp, "c4.vcf.gz"), samples, records, contigs=contigs, compress="bgzf") or os.path.join(tmp, "c4.vcf.gz")
    bcf = write_bcf(os.path.join(tmp, "c4.bcf"), samples, records, contigs=contigs, compress="bgzf") or os.path.join(tmp, "c4.bcf")
    return REAL + derive_scores(tmp, rng, entries, n_synth, keep=0.9, flip=0.25), vcf, bcf, len(keys)


def test_config4_shape_18_score_files(api, tmp_path):
    """18 score files (4 bundled + 14 synthetic, see config4_files) x one cohort, one pass."""
    rng = np.random.default_rng(0x6E696D70)
    paths, vcf, bcf, n_sites = config4_files(str(tmp_path), rng, n=3000)
    assert len(paths) == 18 and n_sites > 700
    for exact in (False, True):
        got = api.run_multi(paths, bcf, exact_order=exact)
        for k, p in enumerate(paths):
            check(got[k], orc.compute_scores_files(p, vcf), exact)


def test_multi_falls_back_when_slab_is_too_small(api, tmp_path, monkeypatch):
    """Rows of all files together exceeding the slab: the files are scored one run each instead."""
    rng = np.random.default_rng(6)
    d = make_dataset(str(tmp_path), rng, n=300, V=90)
    paths = [d["score"]] + derive_scores(str(tmp_path), rng, d["entries"], 2)
    monkeypatch.setenv("NIMPRESS_SLAB_ROWS", "20")
    got = api.run_multi(paths, d["bcf"], exact_order=True)
    for k, p in enumerate(paths):
        assert got[k].rounds > 1
        check(got[k], orc.compute_scores_files(p, d["vcf"]), False)


def test_cli_multi_blocks(api, tmp_path):
    """`nimpress a,b,c genotypes`: per file a "#score" line, then exactly the single-file output."""
    rng = np.random.default_rng(8)
    d = make_dataset(str(tmp_path), rng, n=120, V=60)
    paths = [d["score"]] + derive_scores(str(tmp_path), rng, d["entries"], 2)
    exe = os.path.join(ROOT, "nimpress_b200", "bin", "nimpress")
    multi = subprocess.run([exe, "--exact-order", "--cov=" + d["bed"], ",".join(paths), d["bcf"]], capture_output=True, text=True)
    assert multi.returncode == 0, multi.stderr
    want = ""
    for p in paths:
        one = subprocess.run([exe, "--exact-order", "--cov=" + d["bed"], p, d["bcf"]], capture_output=True, text=True)
        assert one.returncode == 0
        want += f"#score\t{p}\n" + one.stdout
    assert multi.stdout == want
    bad = subprocess.run([exe, paths[0] + "," + str(tmp_path / "nope.score"), d["bcf"]], capture_output=True, text=True)
    assert bad.returncode == 255 and "FATAL Could not open polygenic score file" in bad.stdout


@pytest.mark.parametrize("exact", [False, True], ids=["default", "exact-order"])
def test_resident_multi_c_abi(nb, exact):
    """npc_score_resident_multi over a device slab == npc_reset/score_resident/finish per definition,
    and == the oracle: row lists of different lengths (one empty), shared and private slab rows."""
    rng = np.random.default_rng(21)
    n, V = 30011, 180
    gt = random_cohort(rng, n, V, miss_rate=0.04, n_alt=3)
    lists = [random_rows(rng, V, n_rows=m, n_alt=3) for m in (200, 37, 0, 411, 1)]
    offs = [0.0, -1.5, 2.0, 0.25, 7.0]
    eng = nb.Engine(n, max_rows_per_block=128, n_slots=2)
    eng.set_exact_order(exact)
    cap = eng.resident_reserve(V)
    assert cap >= V
    for r0 in range(0, V, 128):
        slot, view = eng.stage_acquire()
        m = min(128, V - r0)
        view[:m, :gt.shape[1]] = gt[r0:r0 + m].view(np.uint8)
        eng.stage_upload(slot, m, r0)
    got = eng.score_resident_multi(lists, offs)
    for k, rows in enumerate(lists):
        want = orc.score_matrix(gt, n, 2, rows, offset=offs[k])
        sc, nloci, loci = got[k]
        assert nloci == want["nloci"]
        assert_loci_equal(loci, want["loci"])
        a, b = sc, want["scores"]
        assert np.array_equal(np.isnan(a), np.isnan(b))
        ok = np.isfinite(b)
        if exact:
            assert np.array_equal(bits(a[ok]), bits(b[ok]))
        else:
            assert np.all(np.abs(a[ok] - b[ok]) <= 1e-12 * np.maximum(np.abs(b[ok]), 1e-3))
    eng.close()
