"""Test helpers: random cohorts in raw BCF GT encoding, random score rows, parity assertion."""
import numpy as np

import orc

ROW_DTYPE = orc.ROW_DTYPE
_INT = {1: np.int8, 2: np.int16, 4: np.int32}


def random_cohort(rng, n, V, width=1, ploidy=2, miss_rate=0.02, n_alt=1, halfcall_rate=0.002, phased_rate=0.3,
                  sentinel_rate=0.0, invalid_rate=0.0, pad_to=16):
    """gt[V, cols] of dtype int{8,16,32}: (allele+1)<<1|phased, 0|phase = missing allele; row bytes
    padded to a multiple of `pad_to`.  sentinel_rate: fraction of samples given a short call
    (vector_end padding) or a missing-sentinel; invalid_rate: raw negative bytes (invalid BCF)."""
    dt = _INT[width]
    info = np.iinfo(dt)
    cols = n * ploidy
    row_bytes = -(-cols * width // pad_to) * pad_to
    gt = np.zeros((V, row_bytes // width), dtype=dt)
    af = rng.uniform(0.01, 0.5, size=(V, 1, 1))
    alleles = (rng.random((V, n, ploidy)) < af).astype(np.int64)
    if n_alt > 1:
        alleles = alleles * rng.integers(1, n_alt + 1, size=(V, n, ploidy))
    raw = (alleles + 1) << 1
    raw |= (rng.random((V, n, ploidy)) < phased_rate).astype(np.int64)
    miss = rng.random((V, n)) < miss_rate * rng.uniform(0, 2, size=(V, 1))
    raw[miss] &= 1                                    # "./." keeps only the phase bit
    half = rng.random((V, n)) < halfcall_rate
    raw[half, 0] &= 1                                 # "./a"
    if sentinel_rate > 0 and ploidy > 1:
        short = rng.random((V, n)) < sentinel_rate    # haploid call in a wider row
        raw[short, ploidy - 1] = info.min + 1
        msent = rng.random((V, n)) < sentinel_rate / 2
        raw[msent, rng.integers(0, ploidy)] = info.min
        both = rng.random((V, n)) < sentinel_rate / 4  # vector_end first: whole sample ignored
        raw[both, 0] = info.min + 1
    if invalid_rate > 0:
        bad = rng.random((V, n, ploidy)) < invalid_rate
        raw[bad] = rng.integers(-8, 0, size=int(bad.sum()))
    gt[:, :cols] = raw.reshape(V, cols).astype(dt)
    if gt.shape[1] > cols:                            # padding bytes must never be read as samples
        gt[:, cols:] = rng.integers(0, 6, size=(V, gt.shape[1] - cols)).astype(dt)
    return gt


def random_rows(rng, V, n_rows=None, n_alt=1, kinds=(0.8, 0.07, 0.07, 0.06), nan_eaf_rate=0.02, shuffle_gt=True):
    """Score rows in processing order.  kinds = probabilities of GT / NOTCOV / ABSENT / FILTER.
    Several rows may name the same genotype row with different effect alleles."""
    n_rows = V if n_rows is None else n_rows
    rows = np.zeros(n_rows, dtype=ROW_DTYPE)
    rows["kind"] = rng.choice(4, size=n_rows, p=kinds)
    g = rng.integers(0, max(V, 1), size=n_rows) if shuffle_gt else np.arange(n_rows) % max(V, 1)
    rows["gt_row"] = np.where((rows["kind"] == 0) | ((rows["kind"] == 3) & (rng.random(n_rows) < 0.5)), g, -1)   # FILTER rows: with or without a slab row
    rows["ref_is_ea"] = rng.random(n_rows) < 0.27
    rows["eaidx"] = np.where(rows["ref_is_ea"] == 1, 0, rng.integers(1, n_alt + 1, size=n_rows))
    rows["eaidx"] = np.where((rows["gt_row"] < 0) & (rows["kind"] != 3), -1, rows["eaidx"])
    rows["beta"] = np.round(rng.normal(0, 0.05, size=n_rows), 4)
    rows["eaf"] = np.round(rng.uniform(0.01, 0.5, size=n_rows), 4)
    rows["eaf"][rng.random(n_rows) < nan_eaf_rate] = np.nan
    if V == 0:
        rows["kind"] = np.where(rows["kind"] == 0, 2, rows["kind"])
        rows["gt_row"] = -1
    return rows


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def assert_loci_equal(got, want):
    assert len(got) == len(want)
    for f in ("klass", "used", "eaidx", "ngt", "nmiss", "neff"):
        bad = np.nonzero(got[f] != want[f])[0]
        assert bad.size == 0, f"locus field {f} differs at rows {bad[:5]}: {got[f][bad[:5]]} vs {want[f][bad[:5]]}"
    a, b = got["imputed"], want["imputed"]
    assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])


def score_excess(a, b, want, rtol=1e-12):
    """max over finite samples of |a-b| / (rtol*|b| + want["abs_floor"]): <= 1 passes.  Purely relative
    (rtol 1e-12, a thousand times tighter than north_star's 1e-9) plus the absolute floor of
    orc.abs_floor -- 8 * eps * sum|beta| / nloci, the rounding scale of the sum itself."""
    ok = np.isfinite(b)
    if not ok.any():
        return 0.0
    return float((np.abs(a[ok] - b[ok]) / (rtol * np.abs(b[ok]) + want["abs_floor"] + 1e-300)).max())


def assert_parity(got, want, exact=True, rtol=1e-12):
    """got: Engine.finish() dict; want: orc.score_matrix() dict.  exact: scores bit-equal (generic
    kernels, exact-order fused kernel).  Otherwise (default 4-row-tile kernel: same rounded products,
    different association) |a-b| <= rtol*|b| + abs_floor (score_excess) and identical NaN / inf patterns."""
    assert got["nloci"] == want["nloci"], (got["nloci"], want["nloci"])
    assert_loci_equal(got["loci"], want["loci"])
    a, b = got["scores"], want["scores"]
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN pattern differs"
    if exact:
        ok = ~np.isnan(a)
        bad = np.nonzero(bits(a[ok]) != bits(b[ok]))[0]
        assert bad.size == 0, f"{bad.size} scores differ in bits, first {a[ok][bad[:3]]} vs {b[ok][bad[:3]]}"
    else:
        inf = np.isinf(b)
        assert np.array_equal(np.isinf(a), inf) and np.array_equal(a[inf], b[inf]), "inf pattern differs"
        ex = score_excess(a, b, want, rtol)
        assert ex <= 1.0, f"score error is {ex:.3g} x the tolerance (rtol {rtol}, floor {want['abs_floor']:.3g})"
