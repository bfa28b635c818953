"""GPU tests of FORMAT/DS rows.  PARITY UNPINNED: the reference reads GT only
(src/nimpress.nim:384) and lists dosage input under "Future" (README.md:162-165); the expectation is the
oracle's restatement of the reference's per-locus logic on a real-valued raw dosage (oracle/nimpress_oracle.c,
get_raw_dosages_ds / tally_dosages_ds), bit for bit: the tally order is part of the definition."""
import os

import numpy as np
import pytest

import orc
from util_bcf import write_bcf, write_vcf
from util_cohort import assert_parity, random_rows

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nb():
    import __graft_entry__ as g
    g.build()
    import nimpress_b200
    return nimpress_b200


@pytest.fixture(scope="module")
def api():
    import __graft_entry__ as g
    g.build()
    from nimpress_b200 import api
    return api


def random_ds(rng, n, V, miss_rate=0.03):
    stride = -(-n // 32) * 32
    ds = np.zeros((V, stride), dtype=np.float32)
    ds[:, :n] = np.round(rng.uniform(0, 2, size=(V, n)) * rng.choice([0.0, 1.0, 1.0, 1.0], size=(V, 1)), 3)
    bits = ds.view(np.uint32)
    m = rng.random((V, n)) < miss_rate
    kind = rng.integers(0, 3, size=(V, n))
    bits[:, :n][m & (kind == 0)] = 0x7F800001                 # BCF float missing
    bits[:, :n][m & (kind == 1)] = 0x7F800002                 # vector_end
    bits[:, :n][m & (kind == 2)] = 0x7FC00000                 # a NaN
    if V:
        bits[0, :n] = 0x7F800001                              # a row nobody is called at: maxmis, 0/0 tallies
    return ds


def run_ds(nb, ds, n, rows, offset=0.0, policy=None, block_rows=None):
    V = ds.shape[0]
    eng = nb.Engine(n, ploidy=1, gt_width=4, max_rows_per_block=max(len(rows), V, 1), n_slots=2)
    eng.set_dosage_rows(True)
    eng.set_policy(**(policy or {}))
    eng.reset()
    block_rows = block_rows or max(len(rows), 1)
    for r0 in range(0, max(len(rows), 1), block_rows):
        eng.score_host(ds, rows[r0:r0 + block_rows])
    out = eng.finish(offset=offset)
    eng.close()
    return out


@pytest.mark.parametrize("n", [1, 7, 8, 9, 255, 256, 257, 1000, 4099, 70001])
def test_dosage_rows_equal_oracle(nb, n):
    rng = np.random.default_rng(n)
    V = 23
    ds = random_ds(rng, n, V)
    rows = random_rows(rng, V, n_rows=40)
    rows["eaidx"] = np.where(rows["ref_is_ea"] == 1, 0, 1)
    for pol in (dict(), dict(imp_locus="homref", imp_sample="ps", maxmis=0.02), dict(imp_locus="fail", imp_sample="int_fail", mincs=n + 1),
                dict(imp_locus="ignore", imp_missing="ignore", imp_sample="homref", maxmis=1.0)):
        got = run_ds(nb, ds, n, rows, offset=0.25, policy=pol, block_rows=17)
        want = orc.score_matrix(ds, n, 1, rows.astype(orc.ROW_DTYPE), offset=0.25, **pol)
        assert_parity(got, want, exact=True)


def test_dosage_context_requirements(nb):
    e = nb.Engine(100, ploidy=2, gt_width=1, n_slots=0)
    with pytest.raises(nb.NpcError):
        e.set_dosage_rows(True)
    e.close()


def test_dosage_through_the_host(api, tmp_path):
    """nimpress --dosage / nph_params.use_ds on VCF text and BCF: FORMAT/DS is read (after other FORMAT fields, "."
    as missing), REF or ALT as effect allele, absent / FILTER-failed / uncovered rows as for GT; scores, records
    and nloci equal the oracle on the DS matrix, on one device and split over several contexts."""
    rng = np.random.default_rng(12)
    n, V = 53, 40
    samples = [f"s{i}" for i in range(n)]
    ds = random_ds(rng, n, V, miss_rate=0.08)[:, :n].copy()
    recs, ents, rows = [], [], []
    for k in range(V):
        d = ds[k].copy()
        d[d.view(np.uint32) > 0x7F800000] = np.nan
        flt = "FAIL" if k % 9 == 4 else "PASS"
        recs.append(dict(contig="1", pos=1000 + 50 * k, ref="A", alts=["G"], filter=flt, ds=d,
                         gt=((rng.integers(0, 2, size=(n, 2)) + 1) << 1).astype(np.int8)))
        ref_is_ea = k % 3 == 0
        beta, eaf = round(float(rng.normal(0, 0.3)), 4), round(float(rng.uniform(0.05, 0.5)), 4)
        ents.append(("1", 1000 + 50 * k, "A", "A" if ref_is_ea else "G", beta, eaf))
        rows.append((k if flt == "PASS" else -1, 0 if ref_is_ea else 1, beta, eaf, int(ref_is_ea), 0 if flt == "PASS" else 3))
    ents.append(("1", 999999, "A", "G", 0.5, 0.2)); rows.append((-1, -1, 0.5, 0.2, 0, 2))       # absent
    sc = tmp_path / "d.score"
    sc.write_text("x\nd\nc\nhs37d5\n0.5\n" + "\n".join("\t".join(map(str, e)) for e in ents) + "\n")
    rows = np.array(rows, dtype=orc.ROW_DTYPE)
    dsm = np.stack([np.where(np.isfinite(r["ds"]), r["ds"], np.float32(np.nan)) for r in recs]).astype(np.float32)
    want = orc.score_matrix(dsm, n, 1, rows, offset=0.5)
    vcf, bcf = str(tmp_path / "d.vcf.gz"), str(tmp_path / "d.bcf")
    write_vcf(vcf, samples, recs, contigs=["1"], compress="bgzf")
    write_bcf(bcf, samples, recs, ["1"], extra_fmt=True)
    for f in (vcf, bcf):
        got = api.run(str(sc), f, dosage=True, exact_order=True)
        assert got.nloci == want["nloci"] and got.samples == samples
        assert_parity(dict(scores=got.scores, nloci=got.nloci, loci=got.loci), want, exact=True)
    os.environ["NIMPRESS_SPLIT"] = "3"
    try:
        got = api.run(str(sc), bcf, dosage=True)
        assert got.devices == 3
        assert_parity(dict(scores=got.scores, nloci=got.nloci, loci=got.loci), want, exact=False)
    finally:
        del os.environ["NIMPRESS_SPLIT"]
    assert api.main(["--dosage", str(sc), bcf]) == 0
    # GT-only file in dosage mode: malformed input, like a record without GT in the reference
    gt_only = str(tmp_path / "g.bcf")
    write_bcf(gt_only, samples, [{k: v for k, v in r.items() if k != "ds"} for r in recs], ["1"])
    with pytest.raises(api.NimpressInputError):
        api.run(str(sc), gt_only, dosage=True)
