"""GPU tests of the variant-sharded multi-GPU path inside the product boundary:
npc_reduce (one process, one context per GPU, combine over NVLink peer memory), the host library's
`devices` list / `nimpress --devices`, and npc_comm_* (one process per GPU, NCCL loaded at run time).

Expectation: the reference adds every locus in score-file order in one chain (src/nimpress.nim:634-641).
A sharded run adds each contiguous range of score rows in that order (bit for bit in exact-order mode)
and then the ranges' partial sums in range order: per-locus records and nloci are identical, scores differ
from the single chain only by that one re-association (<= 1e-12 relative + orc.abs_floor; contract 1e-9)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import orc
from util_cohort import assert_loci_equal, assert_parity, bits, random_cohort, random_rows, score_excess
from util_files import make_dataset

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def nb():
    import __graft_entry__ as g
    g.build()
    import nimpress_b200 as nb
    return nb


@pytest.fixture(scope="module")
def api():
    import __graft_entry__ as g
    g.build()
    from nimpress_b200 import api
    return api


def n_devices():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("exact", [True, False], ids=["exact-order", "default"])
@pytest.mark.parametrize("D", [2, 3, 5])
def test_reduce_equals_range_order_sum(nb, D, exact):
    """D contexts (spread over the GPUs present, several per GPU when there are fewer) each score a contiguous
    range of the rows; npc_reduce == (p0 + p1) + ... of their npc_partial sums, then npc_normalise, bit for
    bit; every range is bit-equal to the oracle on that range (exact mode); the total agrees with the oracle's
    single chain within the re-association tolerance; records concatenate to the oracle's."""
    rng = np.random.default_rng(100 + D)
    n, V = 4321, 96
    gt = random_cohort(rng, n, V, miss_rate=0.03)
    rows = random_rows(rng, V, n_rows=V + 20)
    ndev = n_devices()
    cuts = [len(rows) * k // D for k in range(D + 1)]
    engines, parts, loci = [], [], []
    for k in range(D):
        e = nb.Engine(n, max_rows_per_block=256, n_slots=2, device=k % ndev)
        e.set_policy()
        e.set_exact_order(exact)
        e.reset()
        e.score_host(gt, rows[cuts[k]:cuts[k + 1]])
        engines.append(e)
    for k, e in enumerate(engines):
        p = e.partial(want_loci=True)
        parts.append(p)
        loci.append(p["loci"])
        want_k = orc.score_matrix(gt, n, 2, rows[cuts[k]:cuts[k + 1]].astype(orc.ROW_DTYPE), offset=0.0)
        got_k = e.finish(offset=0.0)                         # leaves the partial sums in place
        assert_parity(got_k, want_k, exact=exact)
    scores, nloci = nb.reduce_contexts(engines, offset=0.375)
    raw, nloci_raw = nb.reduce_contexts(engines, offset=None)
    total = parts[0]["sums"].copy()
    for p in parts[1:]:
        total = total + p["sums"]                            # IEEE adds in range order, like k_combine
    assert nloci == nloci_raw == sum(p["nloci"] for p in parts)
    ok = ~np.isnan(total)
    assert np.array_equal(np.isnan(raw), ~ok) and np.array_equal(bits(raw[ok]), bits(total[ok]))
    norm = engines[0].normalise(total, nloci, 0.375)
    assert np.array_equal(bits(scores[ok]), bits(norm[ok]))
    want = orc.score_matrix(gt, n, 2, rows.astype(orc.ROW_DTYPE), offset=0.375)
    assert nloci == want["nloci"]
    assert_loci_equal(np.concatenate(loci), want["loci"])
    assert np.array_equal(np.isnan(scores), np.isnan(want["scores"]))
    assert score_excess(scores, want["scores"], want) <= 1.0
    for e in engines:
        e.close()


def test_reduce_rejects_bad_input(nb):
    a, b = nb.Engine(100, n_slots=0), nb.Engine(101, n_slots=0)
    with pytest.raises(nb.NpcError):
        nb.reduce_contexts([a, b])
    with pytest.raises(nb.NpcError):
        nb.reduce_contexts([a, a])
    s, nl = nb.reduce_contexts([a], offset=None)
    assert nl == 0 and not s.any()
    a.close(); b.close()


@pytest.mark.parametrize("exact", [False, True], ids=["default", "exact-order"])
def test_host_devices_split(api, tmp_path, monkeypatch, exact):
    """nph_compute_polygenic_scores with a device list: every GPU present (or, on a one-GPU box, three
    contexts on it via NIMPRESS_SPLIT) scores a contiguous range of the score file; records, nloci and
    WARN text are the oracle's, scores within the re-association tolerance -- over VCF text and BCF, with a
    BED, with rows out of order, with records named by rows of two ranges."""
    rng = np.random.default_rng(77)
    d = make_dataset(str(tmp_path), rng, n=900, V=260, sorted_scores=False)
    ndev = n_devices()
    if ndev >= 2:
        devices = list(range(ndev))
    else:
        devices = None
        monkeypatch.setenv("NIMPRESS_SPLIT", "3")
    for bed in (None, d["bed"]):
        want = orc.compute_scores_files(d["score"], d["vcf"], bed)
        for geno in (d["bcf"], d["vcf"]):
            got = api.run(d["score"], geno, bed, exact_order=exact, devices=devices)
            assert got.devices == (ndev if ndev >= 2 else 3)
            assert got.samples == want["samples"] and got.nloci == want["nloci"] and got.warnings == want["warn"]
            assert_loci_equal(got.loci, want["loci"])
            assert np.array_equal(np.isnan(got.scores), np.isnan(want["scores"]))
            assert score_excess(got.scores, want["scores"], want) <= 1.0


@pytest.mark.parametrize("n_files", [2, 5])
def test_several_score_files_over_several_devices(api, tmp_path, monkeypatch, n_files):
    """nph_compute_polygenic_scores_multi with a device list: every file is cut into ranges of its own, device d scores
    range d of every file in one call (two files: the fused kernel per file; five: the tensor-core contraction), the
    host adds each file's partial sums in device order.  Each result == the oracle on that file alone."""
    from util_files import derive_scores
    rng = np.random.default_rng(80 + n_files)
    d = make_dataset(str(tmp_path), rng, n=600, V=180, sorted_scores=False)
    paths = [d["score"]] + derive_scores(str(tmp_path), rng, d["entries"], n_files - 1)
    ndev = n_devices()
    devices = list(range(ndev)) if ndev >= 2 else None
    if ndev < 2:
        monkeypatch.setenv("NIMPRESS_SPLIT", "3")
    L = api.load_host_library()
    import ctypes as C
    mask = sum(1 << k for k in devices) if devices else 0
    p = api._Params(0, 0, 3, 0, 1, 0, 0, mask, 100, 0.05, 0.001, 0, 0)
    arr = (C.c_char_p * len(paths))(*[os.fsencode(x) for x in paths])
    hs = (C.c_void_p * len(paths))()
    rc = L.nph_compute_polygenic_scores_multi(arr, len(paths), os.fsencode(d["bcf"]), os.fsencode(d["bed"]), C.byref(p), hs)
    assert rc == 0, L.nph_last_error()
    for k, path in enumerate(paths):
        got = api._take_result(L, C.c_void_p(hs[k]))
        want = orc.compute_scores_files(path, d["vcf"], d["bed"])
        assert got.devices == (ndev if ndev >= 2 else 3) and got.nloci == want["nloci"] and got.warnings == want["warn"]
        assert_loci_equal(got.loci, want["loci"])
        assert np.array_equal(np.isnan(got.scores), np.isnan(want["scores"])) and score_excess(got.scores, want["scores"], want) <= 1.0


def test_more_contexts_than_score_rows(api, tmp_path, monkeypatch):
    """Five contexts, three score rows (one of them absent from the file): ranges may be empty; the result is the oracle's."""
    rng = np.random.default_rng(79)
    d = make_dataset(str(tmp_path), rng, n=50, V=30)
    lines = open(d["score"]).read().split("\n")
    small = tmp_path / "small.score"
    small.write_text("\n".join(lines[:5] + lines[5:7] + ["1\t999999999\tA\tC\t0.25\t0.1"]) + "\n")
    want = orc.compute_scores_files(str(small), d["vcf"])
    monkeypatch.setenv("NIMPRESS_SPLIT", "5")
    got = api.run(str(small), d["bcf"])
    assert got.devices == 5 and got.nloci == want["nloci"] and got.warnings == want["warn"]
    assert_loci_equal(got.loci, want["loci"])
    assert np.array_equal(np.isnan(got.scores), np.isnan(want["scores"])) and score_excess(got.scores, want["scores"], want) <= 1.0


def test_cli_devices_option(api, tmp_path, monkeypatch, capfd):
    """`nimpress --devices=<list>` prints what the single-device run prints, to the last digit that the
    re-association leaves alone (compared as floats, 1e-12 relative + floor); bad lists are usage errors."""
    rng = np.random.default_rng(78)
    d = make_dataset(str(tmp_path), rng, n=400, V=120)
    ndev = n_devices()
    if ndev < 2:
        monkeypatch.setenv("NIMPRESS_SPLIT", "4")
    assert api.main(["--devices=0-%d" % max(ndev - 1, 0), d["score"], d["bcf"]]) == 0
    multi = capfd.readouterr().out
    monkeypatch.delenv("NIMPRESS_SPLIT", raising=False)
    assert api.main([d["score"], d["bcf"]]) == 0
    single = capfd.readouterr().out
    want = orc.compute_scores_files(d["score"], d["vcf"])

    def parse(text):
        warn = [ln for ln in text.split("\n") if ln.startswith("WARN")]
        sc = [ln.split("\t") for ln in text.split("\n") if ln and not ln.startswith("WARN")]
        return warn, [s[0] for s in sc], np.array([float(s[1]) for s in sc])
    w1, names1, s1 = parse(multi)
    w2, names2, s2 = parse(single)
    assert w1 == w2 and names1 == names2 == want["samples"]
    assert score_excess(s1, want["scores"], want) <= 1.0 and score_excess(s2, want["scores"], want) <= 1.0
    assert api.main(["--devices=", d["score"], d["bcf"]]) == 1
    assert api.main(["--devices=3-1", d["score"], d["bcf"]]) == 1
    capfd.readouterr()


def test_nccl_combine_two_ranks(tmp_path):
    """npc_comm_init / npc_comm_combine under torchrun with two ranks (needs two GPUs): see tests/nccl_worker.py."""
    if n_devices() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29731", os.path.join(ROOT, "tests", "nccl_worker.py")], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "nccl combine ok" in r.stdout
