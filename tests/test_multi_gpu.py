"""GPU tests of the multi-score path (BASELINE.json configs[3]): several score files over ONE pass of
the genotype file and one resident slab.  The reference has no such mode -- it is one nimpress run
per score file -- so the expectation for file k is the oracle run on file k alone."""
import os
import subprocess

import numpy as np
import pytest

import orc
from util_cohort import assert_loci_equal, bits, random_cohort, random_rows, score_excess
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
        assert score_excess(a, b, want, rtol) <= 1.0
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


def fill_slab(eng, gt, block=128):
    V = gt.shape[0]
    assert eng.resident_reserve(V) >= V
    for r0 in range(0, V, block):
        slot, view = eng.stage_acquire()
        m = min(block, V - r0)
        view[:m, :gt.shape[1]] = gt[r0:r0 + m].view(np.uint8)
        eng.stage_upload(slot, m, r0)


def check_lists(got, gt, n, lists, offs, exact, rtol, threads=1, **pol):
    worst = 0.0
    for k, rows in enumerate(lists):
        want = orc.score_matrix(gt, n, 2, rows, offset=offs[k], threads=threads, **pol)
        sc, nloci, loci = got[k]
        assert nloci == want["nloci"]
        assert_loci_equal(loci, want["loci"])
        a, b = sc, want["scores"]
        assert np.array_equal(np.isnan(a), np.isnan(b)), (k, np.isnan(a).sum(), np.isnan(b).sum())
        ok = np.isfinite(b)
        if exact:
            assert np.array_equal(bits(a[ok]), bits(b[ok]))
        elif ok.any():
            ex = score_excess(a, b, want, rtol)
            worst = max(worst, ex * rtol)
            assert ex <= 1.0, (k, ex)
    return worst


@pytest.mark.parametrize("mode", ["contract", "one-by-one", "exact-order"])
def test_resident_multi_c_abi(nb, mode, monkeypatch):
    """npc_score_resident_multi over a device slab == the oracle per definition: row lists of different
    lengths (one empty), shared and private slab rows, repeated rows, every row kind.  `contract` is the
    tensor-core contraction (1e-9 contract, checked at 1e-12), the others the fused kernel per definition."""
    rng = np.random.default_rng(21)
    n, V = 30011, 180
    gt = random_cohort(rng, n, V, miss_rate=0.04, n_alt=3)
    lists = [random_rows(rng, V, n_rows=m, n_alt=3) for m in (200, 37, 0, 411, 1)]
    offs = [0.0, -1.5, 2.0, 0.25, 7.0]
    if mode == "one-by-one":
        monkeypatch.setenv("NPC_MULTI", "0")
    eng = nb.Engine(n, max_rows_per_block=128, n_slots=2)
    eng.set_exact_order(mode == "exact-order")
    fill_slab(eng, gt)
    got = eng.score_resident_multi(lists, offs)
    assert eng.multi_contractions == (1 if mode == "contract" else 0)
    check_lists(got, gt, n, lists, offs, mode == "exact-order", 1e-12)
    eng.close()


@pytest.mark.parametrize("pol", [dict(), dict(imp_locus="fail", imp_sample="int_fail", mincs=40000, maxmis=0.03),
                                 dict(imp_locus="homref", imp_sample="fail"), dict(imp_locus="ignore", imp_missing="ignore", imp_sample="ps", maxmis=1.0)],
                         ids=["default", "fail-int_fail", "homref-fail", "ignore-ps"])
def test_contraction_policies_and_odd_bytes(nb, pol):
    """The contraction under the policies that poison samples (NaN counter row), with sentinel / short-call /
    invalid bytes in the slab (its scalar decode path), 20 definitions (two launches of <= 16), NaN eaf."""
    rng = np.random.default_rng(33)
    n, V = 7001, 150
    gt = random_cohort(rng, n, V, miss_rate=0.03, n_alt=5, sentinel_rate=0.01, invalid_rate=0.002)
    lists = [random_rows(rng, V, n_rows=int(rng.integers(1, 300)), n_alt=5, nan_eaf_rate=0.05) for _ in range(20)]
    offs = [float(k) for k in range(20)]
    eng = nb.Engine(n, max_rows_per_block=256, n_slots=2)
    eng.set_policy(**pol)
    fill_slab(eng, gt, 256)
    got = eng.score_resident_multi(lists, offs)
    assert eng.multi_contractions == 1
    check_lists(got, gt, n, lists, offs, False, 1e-12, **pol)
    eng.close()


def test_contraction_falls_back_outside_its_range(nb):
    """More than four repeats of one (row, allele) in a definition, or an effect-allele index no int8 code
    reaches: served one by one, same results."""
    rng = np.random.default_rng(34)
    n, V = 3000, 40
    gt = random_cohort(rng, n, V, n_alt=2)
    base = random_rows(rng, V, n_rows=60, n_alt=2, kinds=(1.0, 0, 0, 0))
    rep = base.copy(); rep["gt_row"][:6] = 3; rep["eaidx"][:6] = 1
    far = base.copy(); far["eaidx"][5] = 63
    for lists in ([base, rep, base], [base, far, base]):
        eng = nb.Engine(n, max_rows_per_block=64, n_slots=2)
        fill_slab(eng, gt, 64)
        got = eng.score_resident_multi(lists, [0.0, 1.0, 2.0])
        assert eng.multi_contractions == 0
        check_lists(got, gt, n, lists, [0.0, 1.0, 2.0], False, 1e-12)
        eng.close()


def test_config4_contraction_200k(nb):
    """BASELINE.json configs[3] at its size: 18 definitions (the bundled wood weights + 17 derived) x ~700 loci x
    200,000 samples, one contraction (two launches: 16 + 2 definitions)."""
    offset, ents = read_score(os.path.join(G, "scores", "wood-25282103-height.scores"))
    n, V = 200_000, len(ents)
    rng = np.random.default_rng(0x6E696D70)
    af = np.array([e["eaf"] for e in ents])
    stride = -(-2 * n // 128) * 128
    gt = np.zeros((V, stride), np.int8)
    orc.synth_fill(gt, n, 0, 0x6E696D70, (af * 65536).astype(np.uint32), np.full(V, int(0.005 * (1 << 24)), np.uint32), np.ones(V, np.int32))
    base = np.zeros(V, dtype=orc.ROW_DTYPE)
    base["gt_row"] = np.arange(V)
    base["ref_is_ea"] = [int(e["ref"] == e["ea"]) for e in ents]
    base["eaidx"] = np.where(base["ref_is_ea"] == 1, 0, 1)
    base["beta"] = [e["beta"] for e in ents]
    base["eaf"] = af
    lists, offs = [base], [offset]
    for k in range(17):
        r = base[rng.random(V) < 0.9].copy()
        r["beta"] = np.round(rng.normal(0, 0.05, len(r)), 4)
        flip = rng.random(len(r)) < 0.2
        r["eaidx"][flip] ^= 1; r["ref_is_ea"][flip] ^= 1
        lists.append(r); offs.append(float(k))
    eng = nb.Engine(n, max_rows_per_block=256, n_slots=2)
    fill_slab(eng, gt, 256)
    got = eng.score_resident_multi(lists, offs)
    assert eng.multi_contractions == 1
    worst = check_lists(got, gt, n, lists, offs, False, 1e-12, threads=8)
    print(f"config 4 contraction: worst relative deviation {worst:.3e}")
    eng.close()


def test_multi_wider_layout_restarts_and_scores_one_by_one(api, tmp_path):
    """int16 GT storage and triploid calls: the pass restarts with the wider layout (as for one file) and
    the definitions are scored by the general kernels; results still equal the oracle's per file."""
    rng = np.random.default_rng(9)
    d = make_dataset(str(tmp_path), rng, n=150, V=60, gt_dtype=np.int16, ploidy=3, haploid_rate=0.1)
    paths = [d["score"]] + derive_scores(str(tmp_path), rng, d["entries"], 3)
    got = api.run_multi(paths, d["bcf"], d["bed"])
    for k, p in enumerate(paths):
        check(got[k], orc.compute_scores_files(p, d["vcf"], d["bed"]), False)


def test_multi_single_file_and_errors(api, tmp_path):
    """One file through the multi entry point == the single entry point; a missing score file or genotype
    file is reported like the single call reports it."""
    rng = np.random.default_rng(10)
    d = make_dataset(str(tmp_path), rng, n=90, V=45)
    one = api.run(d["score"], d["bcf"], exact_order=True)
    multi = api.run_multi([d["score"]], d["bcf"], exact_order=True)
    assert len(multi) == 1 and np.array_equal(bits(one.scores), bits(multi[0].scores)) and multi[0].warnings == one.warnings
    with pytest.raises(FileNotFoundError):
        api.run_multi([d["score"], str(tmp_path / "missing.score")], d["bcf"])
    with pytest.raises(FileNotFoundError):
        api.run_multi([d["score"]], str(tmp_path / "missing.bcf"))


@pytest.mark.parametrize("n,V,S", [(1, 3, 3), (5, 1, 4), (255, 64, 3), (257, 65, 19), (513, 130, 40), (515, 1300, 37), (4099, 129, 17)])
def test_contraction_edge_shapes(nb, n, V, S):
    """Tile and k-block edges of the contraction: fewer samples than a tile, one more than a tile, exactly /
    one more than 64 entries, more definitions than a launch holds (several launches)."""
    rng = np.random.default_rng(100 + n)
    gt = random_cohort(rng, n, V, miss_rate=0.05, n_alt=2)
    lists = [random_rows(rng, V, n_rows=int(rng.integers(1, 2 * V + 2)), n_alt=2) for _ in range(S)]
    offs = [0.5 * k for k in range(S)]
    eng = nb.Engine(n, max_rows_per_block=256, n_slots=2)
    fill_slab(eng, gt, 256)
    got = eng.score_resident_multi(lists, offs)
    def repeats(rows):
        g = rows[rows["kind"] == 0]
        return max(np.unique(g["gt_row"].astype(np.int64) * 64 + g["eaidx"], return_counts=True)[1], default=0)
    assert eng.multi_contractions == (1 if max(repeats(r) for r in lists) <= 4 else 0)      # > 4 repeats: one by one
    check_lists(got, gt, n, lists, offs, False, 1e-12)
    # the same slab again with other definitions: the scratch arena and the slab are reused
    lists2 = [random_rows(rng, V, n_rows=V, n_alt=2) for _ in range(3)]
    got2 = eng.score_resident_multi(lists2, [0.0, 0.0, 1.0])
    check_lists(got2, gt, n, lists2, [0.0, 0.0, 1.0], False, 1e-12)
    eng.close()


@pytest.mark.parametrize("mode", ["contract", "one-by-one"])
def test_multi_variant_shards_combine(nb, mode, monkeypatch):
    """Variant-sharded multi-score: two contexts each score their half of every definition's
    rows as raw partial sums (offsets = None); shard.combine_partials adds them in shard order per
    definition, npc_normalise once == the oracle on the whole definitions."""
    import torch
    from nimpress_b200 import shard
    if mode == "one-by-one":
        monkeypatch.setenv("NPC_MULTI", "0")
    rng = np.random.default_rng(55)
    n, V, S = 9001, 120, 5
    gt = random_cohort(rng, n, V, miss_rate=0.03, n_alt=2)
    lists = [random_rows(rng, V, n_rows=int(rng.integers(20, 200)), n_alt=2) for _ in range(S)]
    offs = [0.1 * k for k in range(S)]
    parts = []
    for r in range(2):
        eng = nb.Engine(n, max_rows_per_block=128, n_slots=2)
        fill_slab(eng, gt)
        sub = [rows[slice(*shard.shard_range(len(rows), 2, r))] for rows in lists]
        got = eng.score_resident_multi(sub, None)
        assert eng.multi_contractions == (1 if mode == "contract" else 0)
        parts.append((np.stack([g[0] for g in got]), np.array([g[1] for g in got]), [g[2] for g in got]))
        eng.close()
    total = parts[0][0] + parts[1][0]                       # shard.combine_partials at world size 1 per "rank", added in rank order
    nloci = parts[0][1] + parts[1][1]
    t, nl = shard.combine_partials(torch.from_numpy(total), torch.from_numpy(nloci))
    assert np.array_equal(t.numpy(), total, equal_nan=True) and np.array_equal(nl.numpy(), nloci)
    L = nb.load_library()
    for k in range(S):
        want = orc.score_matrix(gt, n, 2, lists[k], offset=offs[k])
        sc = total[k].copy()
        L.npc_normalise(sc.ctypes.data, n, int(nloci[k]), offs[k])
        assert nloci[k] == want["nloci"]
        assert_loci_equal(np.concatenate([parts[0][2][k], parts[1][2][k]]), want["loci"])
        a, b = sc, want["scores"]
        assert np.array_equal(np.isnan(a), np.isnan(b))
        ok = np.isfinite(b)
        assert score_excess(a, b, want) <= 1.0


@pytest.mark.parametrize("parts", ["2", "3", "7"])
def test_contraction_k_split(nb, parts, monkeypatch):
    """The k-blocks of a tile split over several work units (NPC_MULTI_PARTS forces what the library
    otherwise chooses from the tile count): partial sums added in part order by k_multi_finish, NaN
    counter rows included (imp_sample=fail)."""
    monkeypatch.setenv("NPC_MULTI_PARTS", parts)
    rng = np.random.default_rng(77)
    n, V, S = 2100, 300, 5
    gt = random_cohort(rng, n, V, miss_rate=0.03, n_alt=2)
    lists = [random_rows(rng, V, n_rows=int(rng.integers(200, 600)), n_alt=2) for _ in range(S)]
    offs = [0.25 * k for k in range(S)]
    for pol in (dict(), dict(imp_sample="fail", maxmis=0.1)):
        eng = nb.Engine(n, max_rows_per_block=512, n_slots=2)
        eng.set_policy(**pol)
        fill_slab(eng, gt, 512)
        got = eng.score_resident_multi(lists, offs)
        assert eng.multi_contractions == 1
        check_lists(got, gt, n, lists, offs, False, 1e-12, **pol)
        eng.close()


@pytest.mark.parametrize("S", [1, 2])
def test_contraction_forced_for_one_or_two_definitions(nb, S, monkeypatch):
    """NPC_MULTI=1 sends even one or two definitions through the contraction (by default they take the fused kernel)."""
    monkeypatch.setenv("NPC_MULTI", "1")
    rng = np.random.default_rng(90 + S)
    n, V = 5000, 100
    gt = random_cohort(rng, n, V, miss_rate=0.02, n_alt=2)
    lists = [random_rows(rng, V, n_rows=150, n_alt=2) for _ in range(S)]
    eng = nb.Engine(n, max_rows_per_block=128, n_slots=2)
    fill_slab(eng, gt)
    got = eng.score_resident_multi(lists, [0.5] * S)
    assert eng.multi_contractions == 1
    check_lists(got, gt, n, lists, [0.5] * S, False, 1e-12)
    eng.close()
