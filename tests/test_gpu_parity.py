"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.
Bars: per-locus {class, used, eaidx, ngt, nmiss, neff, imputed} and nloci bit-exact in every mode.
Per-sample scores: bit-exact for the generic kernels and the exact-order fused kernel (same rounded
products, same order as the reference's chain); <= 1e-12 relative for the default 4-row-tile fused
kernel (same products, tile-wise association) -- north_star's bar is 1e-9 relative."""
import itertools
import json
import os

import numpy as np
import pytest

import orc
from util_cohort import random_cohort, random_rows, assert_parity, assert_loci_equal, bits

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def nb():
    import __graft_entry__ as g
    g.build()
    import nimpress_b200
    return nimpress_b200


MODES = ["tile4", "exact"]


def run_engine(nb, gt, n, rows, ploidy=2, offset=0.0, block_rows=None, staged=True, policy=None, max_rows=None, mode="exact"):
    width = gt.dtype.itemsize
    policy = policy or {}
    V = gt.shape[0]
    block_rows = block_rows or max(len(rows), 1)
    eng = nb.Engine(n, ploidy=ploidy, gt_width=width, max_rows_per_block=max_rows or max(block_rows, V, 1),
                    n_slots=2 if staged else 0)
    eng.set_policy(**policy)
    eng.set_exact_order(mode == "exact")
    eng.reset()
    if staged:
        for r0 in range(0, max(len(rows), 1), block_rows):
            eng.score_host(gt, rows[r0:r0 + block_rows])
    else:
        import torch
        d = torch.from_numpy(gt.view(np.uint8).reshape(V, -1) if V else np.zeros((0, 16), np.uint8)).cuda()
        stride = d.shape[1] if V else 16
        for r0 in range(0, max(len(rows), 1), block_rows):
            eng.score_block_device(d if V else None, stride, V, rows[r0:r0 + block_rows])
    out = eng.finish(offset=offset)
    out["launches"] = eng.launches
    eng.close()
    return out


def oracle(gt, n, rows, ploidy=2, offset=0.0, policy=None):
    policy = policy or {}
    return orc.score_matrix(gt, n, ploidy, rows, offset=offset, **policy)


# ---- golden vectors of the reference through the CUDA path ---------------------------------

CASES = json.load(open(os.path.join(G, "set1_expected.json")))["cases"]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("case", CASES, ids=[c["ref"] for c in CASES])
def test_set1_golden_through_gpu(nb, case, mode):
    """tests/test_set1.nim: every expected vector, 1e-4 abs + NaN pattern (checkFloats :14-22),
    and bit-equality with the oracle's file-level run."""
    from util_vcf import build_block
    S1 = os.path.join(G, "set1")
    samples, offset, gt, rows = build_block(os.path.join(S1, "set1.score"), os.path.join(S1, "set1.vcf.gz"),
                                            os.path.join(S1, "set1.bed") if case["cov"] else None,
                                            ignorefilt=case["ignorefilt"])
    pol = dict(imp_locus=case["imp_locus"], imp_missing=case["imp_missing"], imp_sample=case["imp_sample"],
               maxmis=case["maxmis"], mincs=case["mincs"])
    got = run_engine(nb, gt, len(samples), rows, offset=offset, policy=pol, mode=mode)
    exp = np.array([np.nan if e is None else (e[1] - e[2] if isinstance(e, list) else e) for e in case["expected"]])
    assert np.array_equal(np.isnan(got["scores"]), np.isnan(exp))
    ok = ~np.isnan(exp)
    assert np.all(np.abs(got["scores"][ok] - exp[ok]) <= 1e-4)
    ref = orc.compute_scores_files(os.path.join(S1, "set1.score"), os.path.join(S1, "set1.vcf.gz"),
                                   os.path.join(S1, "set1.bed") if case["cov"] else None,
                                   afmisp=case["afmisp"], ignorefilt=case["ignorefilt"], **pol)
    assert_parity(got, ref, exact=mode == "exact")


# ---- randomised cohorts ----------------------------------------------------------------------

@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("n", [1, 7, 8, 9, 255, 256, 257, 1000, 4099, 100003])
def test_sample_count_edges_i8(nb, n, mode):
    rng = np.random.default_rng(n)
    V = 37
    gt = random_cohort(rng, n, V, miss_rate=0.03)
    rows = random_rows(rng, V, n_rows=50)
    assert_parity(run_engine(nb, gt, n, rows, offset=0.5, mode=mode), oracle(gt, n, rows, offset=0.5), exact=mode == "exact")


@pytest.mark.parametrize("width,ploidy", [(1, 1), (1, 2), (1, 3), (2, 2), (2, 4), (4, 2), (4, 1)])
def test_widths_and_ploidies(nb, width, ploidy):
    rng = np.random.default_rng(100 * width + ploidy)
    n, V = 3001, 40
    gt = random_cohort(rng, n, V, width=width, ploidy=ploidy, miss_rate=0.03, n_alt=3, sentinel_rate=0.02)
    rows = random_rows(rng, V, n_rows=60, n_alt=3)
    assert_parity(run_engine(nb, gt, n, rows, ploidy=ploidy), oracle(gt, n, rows, ploidy=ploidy))


@pytest.mark.parametrize("mode", MODES)
def test_slow_path_alleles_sentinels_invalid(nb, mode):
    """int8 diploid rows with alleles >= 7 (bytes >= 16), vector_end / missing sentinels and raw
    negative bytes: every chunk that is not 'all bytes < 16' takes the exact slow decode."""
    rng = np.random.default_rng(5)
    n, V = 5003, 48
    gt = random_cohort(rng, n, V, miss_rate=0.04, n_alt=12, sentinel_rate=0.01, invalid_rate=0.001)
    rows = random_rows(rng, V, n_rows=96, n_alt=12)
    assert_parity(run_engine(nb, gt, n, rows, mode=mode), oracle(gt, n, rows), exact=mode == "exact")


POLICIES = [dict(imp_locus=l, imp_missing=m, imp_sample=s, maxmis=x, mincs=c)
            for l, m, s, x, c in itertools.product(("ps", "homref", "fail", "ignore"), ("homref", "ignore"),
                                                   ("ps", "homref", "fail", "int_ps", "int_fail"),
                                                   (0.0, 0.03, 1.0), (0, 1900))]


@pytest.mark.parametrize("mode", MODES)
def test_policy_grid(nb, mode):
    """Every --imp-locus x --imp-missing x --imp-sample x --maxmis x --mincs combination on one
    cohort whose per-locus missing rates straddle the thresholds."""
    rng = np.random.default_rng(11)
    n, V = 2000, 60
    gt = random_cohort(rng, n, V, miss_rate=0.03)
    rows = random_rows(rng, V, n_rows=80)
    for pol in POLICIES:
        assert_parity(run_engine(nb, gt, n, rows, offset=-0.125, policy=pol, mode=mode), oracle(gt, n, rows, offset=-0.125, policy=pol),
                      exact=mode == "exact")


@pytest.mark.parametrize("mode", MODES)
def test_maxmis_boundary_exact(nb, mode):
    """nmissing/n > maxmis is a strict fp64 compare (src/nimpress.nim:565-566): rows with exactly
    k missing of n against thresholds k/n, nextafter(k/n) on both sides."""
    n = 1000
    rng = np.random.default_rng(3)
    ks = [0, 1, 49, 50, 51, 333, 1000]
    gt = random_cohort(rng, n, len(ks), miss_rate=0.0, halfcall_rate=0.0)
    for i, k in enumerate(ks):
        gt[i, :2 * k] = 0
    rows = random_rows(rng, len(ks), n_rows=len(ks), kinds=(1, 0, 0, 0), shuffle_gt=False)
    for k in (1, 50, 333):
        for thr in (k / n, np.nextafter(k / n, 0), np.nextafter(k / n, 1)):
            pol = dict(maxmis=float(thr), imp_sample="int_ps", mincs=0)
            assert_parity(run_engine(nb, gt, n, rows, policy=pol, mode=mode), oracle(gt, n, rows, policy=pol), exact=mode == "exact")


@pytest.mark.parametrize("mode", MODES)
def test_empty_and_degenerate_blocks(nb, mode):
    rng = np.random.default_rng(1)
    n = 100
    gt = random_cohort(rng, n, 4)
    # no rows at all: nloci = 0 -> 0/0 = NaN for everyone (reference: scores /= 0*2)
    ex = mode == "exact"
    got = run_engine(nb, gt, n, np.zeros(0, dtype=orc.ROW_DTYPE), offset=1.0, mode=mode)
    assert got["nloci"] == 0 and np.all(np.isnan(got["scores"]))
    # only constant rows, no genotype slab
    rows = random_rows(rng, 0, n_rows=9)
    assert_parity(run_engine(nb, np.zeros((0, 16), np.int8), n, rows, mode=mode), oracle(np.zeros((1, 16), np.int8), n, rows), exact=ex)
    # all rows ignored
    pol = dict(imp_locus="ignore", imp_missing="ignore", maxmis=0.0)
    gt2 = random_cohort(rng, n, 4, miss_rate=0.5)
    rows = random_rows(rng, 4, n_rows=8)
    assert_parity(run_engine(nb, gt2, n, rows, policy=pol, mode=mode), oracle(gt2, n, rows, policy=pol), exact=ex)
    # beta = 0, inf and NaN, eaf NaN: NaN*0 = NaN must poison exactly the reference's samples
    rows = random_rows(rng, 4, n_rows=12)
    rows["beta"][:4] = [0.0, np.inf, np.nan, -0.0]
    rows["eaf"][4:6] = np.nan
    for pol in (dict(imp_sample="fail"), dict(imp_sample="ps"), dict(imp_locus="fail", maxmis=0.0)):
        assert_parity(run_engine(nb, gt2, n, rows, policy=pol, mode=mode), oracle(gt2, n, rows, policy=pol), exact=ex)


@pytest.mark.parametrize("mode", MODES)
def test_streaming_blocks_equal_single_block(nb, mode):
    """Rows split over many npc_score_block calls (pinned ring, 3 in flight) give the same bits as
    one block, and the device-resident entry point gives the same bits as the staged one."""
    rng = np.random.default_rng(21)
    n, V = 20011, 300
    gt = random_cohort(rng, n, V, miss_rate=0.02)
    rows = random_rows(rng, V, n_rows=300)
    want = oracle(gt, n, rows)
    one = run_engine(nb, gt, n, rows, mode=mode)
    many = run_engine(nb, gt, n, rows, block_rows=17, max_rows=300, mode=mode)
    dev = run_engine(nb, gt, n, rows, staged=False, block_rows=64, max_rows=300, mode=mode)
    for got in (one, many, dev):
        assert_parity(got, want, exact=mode == "exact")


def test_split_count_accumulate_matches_fused(nb):
    """npc_count_block_device + npc_accumulate_block_device (sample-sharded form) on two sample
    slabs with summed integer counts == the whole cohort on one context, bit for bit."""
    import torch
    rng = np.random.default_rng(33)
    n, V = 6000, 64
    gt = random_cohort(rng, n, V, miss_rate=0.04)
    rows = random_rows(rng, V, n_rows=90)
    want = oracle(gt, n, rows, offset=0.1)
    cut = 2504                                        # slab boundary, multiple of 8
    slabs = [np.ascontiguousarray(gt[:, :2 * cut]), np.ascontiguousarray(gt[:, 2 * cut:2 * n])]
    engs, dev, counts = [], [], []
    for s in slabs:
        ns = s.shape[1] // 2
        pad = -(-s.shape[1] // 16) * 16
        sp = np.zeros((V, pad), np.int8); sp[:, :s.shape[1]] = s
        e = nb.Engine(ns, max_rows_per_block=128)
        e.set_cohort_size(n); e.reset()
        d = torch.from_numpy(sp.view(np.uint8)).cuda()
        c = torch.zeros((len(rows), 2), dtype=torch.int64, device="cuda")
        e.count_block_device(d, pad, V, rows, c)
        engs.append(e); dev.append((d, pad)); counts.append(c)
    torch.cuda.synchronize()
    total = counts[0] + counts[1]                     # the all-reduce of a sharded run
    scores = []
    for e, (d, pad) in zip(engs, dev):
        e.accumulate_block_device(d, pad, V, rows, total)
        out = e.finish(offset=0.1)
        scores.append(out["scores"])
        assert_loci_equal(out["loci"], want["loci"])
        assert out["nloci"] == want["nloci"]
        e.close()
    got = np.concatenate(scores)
    assert np.array_equal(np.isnan(got), np.isnan(want["scores"]))
    ok = ~np.isnan(got)
    assert np.array_equal(bits(got[ok]), bits(want["scores"][ok]))


def test_synth_generator_matches_oracle(nb):
    import torch
    rng = np.random.default_rng(9)
    n, V = 10007, 33
    af = rng.integers(600, 32768, size=V).astype(np.uint32)
    ms = rng.integers(0, 1 << 20, size=V).astype(np.uint32)
    alt = rng.integers(1, 4, size=V).astype(np.int32)
    stride = -(-2 * n // 128) * 128
    host = np.zeros((V, stride), np.int8)
    orc.synth_fill(host, n, 1000, 0x6E696D70, af, ms, alt)
    eng = nb.Engine(n, max_rows_per_block=64)
    d = torch.zeros((V, stride), dtype=torch.int8, device="cuda")
    eng.synth_fill_device(d, stride, 1000, V, 0x6E696D70, torch.from_numpy(af.view(np.int32)).cuda(),
                          torch.from_numpy(ms.view(np.int32)).cuda(), torch.from_numpy(alt).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(d.cpu().numpy()[:, :2 * n], host[:, :2 * n])
    eng.close()


@pytest.mark.parametrize("mode", MODES)
def test_config2_shape_wood_100k(nb, mode):
    """BASELINE.json configs[1]: the 697 wood-height loci x 100,000 synthetic samples, 0.5% missing,
    default policies; whole result bit-equal to the oracle."""
    from util_vcf import read_score
    offset, ents = read_score(os.path.join(G, "scores", "wood-25282103-height.scores"))
    n, V = 100_000, len(ents)
    rng = np.random.default_rng(0x6E696D70)
    af = np.array([e["eaf"] for e in ents])
    stride = -(-2 * n // 128) * 128
    gt = np.zeros((V, stride), np.int8)
    orc.synth_fill(gt, n, 0, 0x6E696D70, (af * 65536).astype(np.uint32), np.full(V, int(0.005 * (1 << 24)), np.uint32),
                   np.ones(V, np.int32))
    rows = np.zeros(V, dtype=orc.ROW_DTYPE)
    rows["gt_row"] = np.arange(V)
    rows["ref_is_ea"] = [int(e["ref"] == e["ea"]) for e in ents]
    rows["eaidx"] = np.where(rows["ref_is_ea"] == 1, 0, 1)
    rows["beta"] = [e["beta"] for e in ents]
    rows["eaf"] = af
    assert V == 697
    assert_parity(run_engine(nb, gt, n, rows, offset=offset, staged=False, mode=mode), oracle(gt, n, rows, offset=offset),
                  exact=mode == "exact")


@pytest.mark.parametrize("env", [dict(NPC_FUSED="0"), dict(NPC_EXACT="1"),
                                 dict(NPC_EXACT="1", NPC_TILE_SR="2", NPC_TILE_SC="10", NPC_TILE_A="1"),
                                 dict(NPC_EXACT="1", NPC_TILE_K="2", NPC_TILE_SR="5", NPC_TILE_SC="13", NPC_TILE_L="9"),
                                 dict(NPC_TILE_SR="2", NPC_TILE_SC="10", NPC_TILE_A="1"),
                                 dict(NPC_TILE_SR="5", NPC_TILE_SC="13", NPC_TILE_L="9", NPC_TILE_A="2"), dict(NPC_TILE_K="2"),
                                 dict(NPC_TILE_GR="1"), dict(NPC_TILE_GR="3"), dict(NPC_TILE_GR="4", NPC_TILE_K="2")],
                         ids=lambda e: ",".join(f"{k[4:]}={v}" for k, v in e.items()))
def test_kernel_paths_agree(nb, env, monkeypatch):
    """Every kernel path and launch shape gives the oracle's per-locus records; the exact-order paths
    (two-kernel sequence, tile kernel in exact mode under several ring shapes) give its bits, the
    default tile kernel agrees to 1e-12 under every ring shape and every sample-slab x row-group grid."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    rng = np.random.default_rng(77)
    n, V = 40009, 203
    gt = random_cohort(rng, n, V, miss_rate=0.03, n_alt=9, sentinel_rate=0.003)
    rows = random_rows(rng, V, n_rows=260, n_alt=9)
    want = oracle(gt, n, rows, offset=0.5)
    eng = nb.Engine(n, max_rows_per_block=512)
    kind = eng.kernel_shape["fused"]
    eng.close()
    exact = "NPC_EXACT" in env or env.get("NPC_FUSED") == "0"
    assert kind == (0 if env.get("NPC_FUSED") == "0" else 1 if exact else 2)
    mode = None                                        # leave the context's default (set by the environment)
    for kw in (dict(staged=False, max_rows=512), dict(staged=True, block_rows=37, max_rows=512)):
        V_ = gt.shape[0]
        eng = nb.Engine(n, max_rows_per_block=512, n_slots=2 if kw["staged"] else 0)
        eng.set_policy(); eng.reset()
        if kw["staged"]:
            for r0 in range(0, len(rows), 37):
                eng.score_host(gt, rows[r0:r0 + 37])
        else:
            import torch
            d = torch.from_numpy(gt.view(np.uint8).reshape(V_, -1)).cuda()
            eng.score_block_device(d, d.shape[1], V_, rows)
        got = eng.finish(offset=0.5)
        eng.close()
        assert_parity(got, want, exact=exact)


@pytest.mark.parametrize("mode", MODES)
def test_full_sample_width_500k(nb, mode):
    """BASELINE.json configs[2] at its full sample width (500,000 samples, the launch shape bench.py
    times: 148 CTAs x 14 consumer warps) on 1,024 variants of the synthetic cohort generated ON THE
    DEVICE (npc_synth_fill_device), against the oracle on the host-generated copy of the same cohort."""
    import torch
    n, V, seed = 500_000, 1024, 0x6E696D70
    rng = np.random.default_rng(seed)
    af = rng.uniform(0.01, 0.5, size=V)
    af_thr = (af * 65536).astype(np.uint32)
    ms_thr = (rng.uniform(0, 0.1, size=V) * (1 << 24)).astype(np.uint32)      # per-locus missing rate U(0, 10%): half exceed --maxmis
    alt = np.ones(V, np.int32)
    stride = -(-2 * n // 128) * 128
    host = np.zeros((V, stride), np.int8)
    orc.synth_fill(host, n, 0, seed, af_thr, ms_thr, alt)
    rows = random_rows(rng, V, n_rows=V, kinds=(0.85, 0.05, 0.05, 0.05), shuffle_gt=False)
    want = oracle(host, n, rows, offset=0.0)
    eng = nb.Engine(n, max_rows_per_block=V)
    eng.set_policy(); eng.set_exact_order(mode == "exact"); eng.reset()
    d = torch.empty((V, stride), dtype=torch.uint8, device="cuda")
    eng.synth_fill_device(d, stride, 0, V, seed, torch.from_numpy(af_thr.view(np.int32)).cuda(),
                          torch.from_numpy(ms_thr.view(np.int32)).cuda(), torch.from_numpy(alt).cuda())
    eng.score_block_device(d, stride, V, rows)
    got = eng.finish()
    shape = eng.kernel_shape
    eng.close()
    assert shape["grid"] == 148 and shape["fused"] == (1 if mode == "exact" else 2)
    assert (want["loci"]["klass"] == orc.CLASS_MAXMIS).sum() > 100 and (want["loci"]["klass"] == orc.CLASS_OK).sum() > 100
    assert_parity(got, want, exact=mode == "exact")


@pytest.mark.parametrize("imp_locus,maxmis", [("ps", 0.05), ("homref", 0.02), ("fail", 1.0)])
def test_config5_policies_50k(nb, imp_locus, maxmis):
    """BASELINE.json configs[4]: 50,000 samples x 10,000 loci, per-locus missing rate U(0, 10%), 10% of
    records FILTER-failed, 5% of score rows absent, 30% not covered; --imp-locus x --maxmis with the
    default --imp-sample int_ps.  Per-locus records bit-equal, scores within 1e-12 (default kernel)."""
    n, V, seed = 50_000, 10_000, 0x6E696D70
    rng = np.random.default_rng(seed + 5)
    af = rng.uniform(0.01, 0.5, size=V)
    stride = -(-2 * n // 128) * 128
    host = np.zeros((V, stride), np.int8)
    orc.synth_fill(host, n, 0, seed, (af * 65536).astype(np.uint32), (rng.uniform(0, 0.1, size=V) * (1 << 24)).astype(np.uint32),
                   np.ones(V, np.int32))
    rows = random_rows(rng, V, n_rows=V, kinds=(0.55, 0.30, 0.05, 0.10), shuffle_gt=False)
    pol = dict(imp_locus=imp_locus, maxmis=maxmis)
    want = oracle(host, n, rows, policy=pol)
    got = run_engine(nb, host, n, rows, policy=pol, staged=False, mode="tile4")
    assert_parity(got, want, exact=False)


def test_c_abi_error_paths(nb):
    """The C ABI refuses bad calls with the documented codes instead of computing something."""
    import ctypes as C
    import torch
    L = nb.load_library()
    h = C.c_void_p()
    assert L.npc_create(C.byref(h), 0, 100, 2, 3, 16, 2) == -1 and b"invalid" in L.npc_last_error(None)      # gt_width 3
    assert L.npc_create(C.byref(h), 0, 100, 2, 1, 16, 1) == -1                                                # one staging slot
    assert L.npc_create(C.byref(h), 99, 100, 2, 1, 16, 2) == -1                                               # no such device
    eng = nb.Engine(100, max_rows_per_block=16, n_slots=2)
    rows = np.zeros(4, dtype=nb.ROW_DTYPE)
    assert L.npc_score_block(eng.h, 0, 1, rows.ctypes.data, 4) == -4                                          # slot not acquired
    slot, view = eng.stage_acquire()
    assert L.npc_score_block(eng.h, slot, 17, rows.ctypes.data, 4) == -1                                      # more rows than the block holds
    big = np.zeros(17, dtype=nb.ROW_DTYPE)
    assert L.npc_score_block(eng.h, slot, 1, big.ctypes.data, 17) == -1
    eng.score_block(slot, 1, rows[:1])
    s2, _ = eng.stage_acquire()
    with pytest.raises(nb.NpcError):                                                                         # ring exhausted: both lent / in flight
        eng.stage_acquire(); eng.stage_acquire()
    d = torch.zeros(64, dtype=torch.uint8, device="cuda")
    assert L.npc_score_block_device(eng.h, d.data_ptr() + 1, 208, 1, rows.ctypes.data, 1, 0) == -1            # misaligned slab
    assert L.npc_score_block_device(eng.h, d.data_ptr(), 100, 1, rows.ctypes.data, 1, 0) == -1                # stride % 16
    assert L.npc_score_block_device(eng.h, d.data_ptr(), 64, 1, rows.ctypes.data, 1, 0) == -1                 # stride < row bytes
    assert L.npc_stage_upload(eng.h, s2, 1, 0) == -1                                                          # no resident slab reserved
    bad = nb.cuda._Policy(7, 0, 0, 0, 0, 0.0)
    assert L.npc_set_policy(eng.h, C.byref(bad)) == -1
    eng.close()
    e0 = nb.Engine(100, max_rows_per_block=16, n_slots=0)
    with pytest.raises(nb.NpcError, match="staging ring"):
        e0.stage_acquire()
    e0.close()


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("n", [9, 4099, 100003])
def test_int16_rows_through_the_fused_kernel(nb, n, mode):
    """int16 GT storage (records with more than 63 alleles; the reference treats every width alike,
    src/nimpress.nim:381-391): diploid int16 cohorts run the fused pair kernel too -- a sample is one 32-bit word,
    two samples per table lookup -- with alleles beyond ALT2, sentinels and the cohort's partial last chunk on the
    exact decode.  Parity unpinned by the reference (no int16 fixture); the oracle widens like htslib."""
    rng = np.random.default_rng(1600 + n)
    V = 44
    gt = random_cohort(rng, n, V, width=2, miss_rate=0.03, n_alt=2, sentinel_rate=0.003)
    gt[5:9] = random_cohort(rng, n, 4, width=2, miss_rate=0.02, n_alt=90, sentinel_rate=0.01)[:, :gt.shape[1]]   # allele numbers an int8 cannot hold
    rows = random_rows(rng, V, n_rows=70, n_alt=3)
    eng = nb.Engine(n, ploidy=2, gt_width=2, max_rows_per_block=128, n_slots=2)
    eng.set_policy(); eng.set_exact_order(mode == "exact"); eng.reset()
    eng.score_host(gt, rows)
    got = eng.finish(offset=0.5)
    shape = eng.kernel_shape
    eng.close()
    assert shape["fused"] == (1 if mode == "exact" else 2), shape          # not the two-kernel sequence
    assert_parity(got, oracle(gt, n, rows, offset=0.5), exact=mode == "exact")


def test_int16_full_sample_width_500k(nb):
    """The bench shard's sample width in int16 storage: the int8 synthetic cohort widened on the host, 256 variants."""
    import torch
    n, V, seed = 500_000, 256, 0x6E696D70
    rng = np.random.default_rng(seed + 16)
    af_thr = (rng.uniform(0.01, 0.5, size=V) * 65536).astype(np.uint32)
    ms_thr = (rng.uniform(0, 0.1, size=V) * (1 << 24)).astype(np.uint32)
    stride8 = -(-2 * n // 128) * 128
    host8 = np.zeros((V, stride8), np.int8)
    orc.synth_fill(host8, n, 0, seed, af_thr, ms_thr, np.ones(V, np.int32))
    host16 = host8.astype(np.int16)
    rows = random_rows(rng, V, n_rows=V, kinds=(0.9, 0.04, 0.03, 0.03), shuffle_gt=False)
    want = oracle(host16, n, rows)
    eng = nb.Engine(n, ploidy=2, gt_width=2, max_rows_per_block=V)
    eng.set_policy(); eng.reset()
    d = torch.from_numpy(host16.view(np.uint8).reshape(V, -1)).cuda()
    eng.score_block_device(d, d.shape[1], V, rows)
    got = eng.finish()
    shape = eng.kernel_shape
    eng.close()
    assert shape["fused"] == 2 and shape["grid"] == 148, shape
    assert_parity(got, want, exact=False)


@pytest.mark.parametrize("mode", MODES)
def test_pair_table_every_code_combination(nb, mode):
    """The pair lookup of the fused kernel (npc_fused5.cuh) indexes a table by the four allele codes of two samples:
    every one of the 4^4 combinations of {missing, REF, ALT1, ALT2} -- with either phase bit -- in every word position
    of a chunk, for effect alleles REF, ALT1, ALT2 (a table each), ALT3 and ALT7 (the no-match table) and in both row
    parities of a tile, against the oracle."""
    rng = np.random.default_rng(44)
    combos = np.array([[a, b, c, d] for a in range(4) for b in range(4) for c in range(4) for d in range(4)], dtype=np.int64)   # [256, 4]
    V = 24
    n = 2 * 256 * 4                                            # every combination in each of the 4 word positions
    gt = np.zeros((V, 2 * n), dtype=np.int8)
    for v in range(V):
        perm = np.concatenate([rng.permutation(256) for _ in range(4)])
        codes = combos[perm].reshape(-1)                       # allele codes of consecutive samples, 2 per sample
        raw = (codes << 1) | rng.integers(0, 2, size=codes.size)
        gt[v] = np.roll(raw.reshape(-1, 4), v % 4, axis=0).reshape(-1).astype(np.int8)     # shift word positions row by row
    rows = np.zeros(40, dtype=nb.ROW_DTYPE)
    rows["gt_row"] = rng.integers(0, V, size=40)
    rows["eaidx"] = np.tile([0, 1, 2, 3, 7], 8)
    rows["beta"] = np.round(rng.normal(0, 0.3, 40), 4)
    rows["eaf"] = np.round(rng.uniform(0.05, 0.5, 40), 4)
    rows["ref_is_ea"] = rows["eaidx"] == 0
    got = run_engine(nb, gt, n, rows, offset=0.0, policy=dict(maxmis=1.0), mode=mode)
    assert_parity(got, oracle(gt, n, rows, policy=dict(maxmis=1.0)), exact=mode == "exact")


@pytest.mark.parametrize("n", [850_000, 1_150_000, 1_300_000, 2_000_000])
@pytest.mark.parametrize("mode", MODES)
def test_wide_cohorts(nb, mode, n):
    """850,000 samples: one chunk per thread on 23 consumer warps (the 72-register instance of the kernel).  1,150,000: two
    chunk sets per tile on 16 warps, each its own raw stage, deciders on groups of 4 tiles (the rings leave a short lag).  1,300,000 / 2,000,000 samples on one GPU:
    more than the tile kernel keeps resident (~1.2 M) -- the rows are tallied and
    decided over all samples first, then the tile kernel runs in "decided" mode once per slab of the sample axis
    (kernel_shape fused = 3): per-locus records equal, scores bit-equal in exact-order mode."""
    import torch
    V, seed = 48, 0x6E696D70
    rng = np.random.default_rng(seed + 3)
    af_thr = (rng.uniform(0.01, 0.5, size=V) * 65536).astype(np.uint32)
    ms_thr = (rng.uniform(0, 0.1, size=V) * (1 << 24)).astype(np.uint32)
    alt = np.ones(V, np.int32)
    stride = -(-2 * n // 128) * 128
    host = np.zeros((V, stride), np.int8)
    orc.synth_fill(host, n, 0, seed, af_thr, ms_thr, alt)
    rows = random_rows(rng, V, n_rows=V + 9, kinds=(0.85, 0.05, 0.05, 0.05))
    want = orc.score_matrix(host, n, 2, rows.astype(orc.ROW_DTYPE), offset=0.25, threads=8)
    eng = nb.Engine(n, max_rows_per_block=128)
    eng.set_policy(); eng.set_exact_order(mode == "exact"); eng.reset()
    d = torch.from_numpy(host.view(np.uint8)).cuda()
    eng.score_block_device(d, stride, V, rows)
    got = eng.finish(offset=0.25)
    shape = eng.kernel_shape
    eng.close()
    assert shape["fused"] == (3 if n > 1_200_000 else 1 if mode == "exact" else 2), shape
    if n == 850_000: assert (shape["chunks_per_thread"], shape["consumer_warps"]) == (1, 23), shape
    if n == 1_150_000: assert (shape["chunks_per_thread"], shape["consumer_warps"], shape["decider_tiles"]) == (2, 16, 4), shape
    # threads=8 in the oracle adds per-thread partial sums: compare the scores at the re-association tolerance, the records exactly
    assert_parity(got, want, exact=False)
    if mode == "exact":
        # bit-equality of the chain: samples whose imputed values do not depend on the cohort tallies
        pol = dict(maxmis=1.0, imp_sample="ps")
        a = orc.score_matrix(host[:, :2 * 4096].copy(), 4096, 2, rows.astype(orc.ROW_DTYPE), offset=0.25, **pol)
        eng = nb.Engine(n, max_rows_per_block=128)
        eng.set_policy(**pol); eng.set_exact_order(True); eng.reset()
        eng.score_block_device(d, stride, V, rows)
        b = eng.finish(offset=0.25)
        eng.close()
        ok = ~np.isnan(a["scores"])
        assert np.array_equal(bits(a["scores"][ok]), bits(b["scores"][:4096][ok]))


def test_long_launch_split(nb, monkeypatch):
    """A long launch (>= 1.5 GB of genotypes, or >= 192 MB where short launches split the rows too) may split the grid
    differently from a short one (more row groups, better-filled warps: 250,000 samples run 74 slabs x 2 row groups on 14 warps when short, 49 x 3 on 20 warps when long).  With the threshold
    lowered to 1 MB the same context serves a short and a "long" launch; both match the oracle, the records exactly."""
    rng = np.random.default_rng(77)
    n, V = 250_000, 256
    gt = random_cohort(rng, n, V, miss_rate=0.01, n_alt=2)
    rows = random_rows(rng, V, n_rows=V + 40, n_alt=2)
    want = oracle(gt, n, rows)
    eng = nb.Engine(n, max_rows_per_block=512, n_slots=2)
    short, long_ = eng.kernel_shape_for(16), eng.kernel_shape_for(1 << 20)
    assert short["fused"] == long_["fused"] == 2 and long_["row_groups"] > short["row_groups"], (short, long_)
    for mb in ("1", "1000000"):
        monkeypatch.setenv("NPC_TILE_LONG_MB", mb)
        eng.set_policy(); eng.reset()
        eng.score_host(gt, rows)
        assert_parity(eng.finish(), want, exact=False)
    eng.close()


def test_kernel_shape_choice_and_contexts_side_by_side(nb):
    """Two findings of the round-2 fuzz (tools/fuzz_parity.py).  157,929 samples: the best-scoring split of the grid (7 row
    groups) does not fit shared memory -- the next candidate must be taken, not the slow generic path.  And the
    dynamic-shared-memory attribute belongs to the kernel function, not to a context: a second context with a smaller
    shape must not shrink what the first one may launch."""
    rng = np.random.default_rng(58)
    n, V = 157_929, 40
    gt = random_cohort(rng, n, V, miss_rate=0.01, n_alt=2)
    rows = random_rows(rng, V, n_rows=52, n_alt=2)
    want = oracle(gt, n, rows)
    for exact in (True, False):
        big = nb.Engine(n, max_rows_per_block=64, n_slots=2)
        big.set_policy(); big.set_exact_order(exact); big.reset()
        assert big.kernel_shape["fused"] == (1 if exact else 2), big.kernel_shape
        small_gt = random_cohort(rng, 4000, 8)
        small_rows = random_rows(rng, 8, n_rows=8)
        small = nb.Engine(4000, max_rows_per_block=64, n_slots=2)       # created AFTER big: sets its own, smaller shapes
        small.set_policy(); small.set_exact_order(exact); small.reset()
        small.score_host(small_gt, small_rows)
        big.score_host(gt, rows)
        assert_parity(big.finish(), want, exact=exact)
        assert_parity(small.finish(), oracle(small_gt, 4000, small_rows), exact=exact)
        big.close(); small.close()
