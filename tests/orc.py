"""ctypes binding of the CPU oracle (oracle/nimpress_oracle.h).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "oracle", "_build", "libnimpress_oracle.so")

LOCUS = {"ps": 0, "homref": 1, "fail": 2, "ignore": 3}
MISSING = {"homref": 0, "ignore": 1}
SAMPLE = {"ps": 0, "homref": 1, "fail": 2, "int_ps": 3, "int_fail": 4}
CLASS_OK, CLASS_NOTCOV, CLASS_ABSENT, CLASS_FILTER, CLASS_MAXMIS = range(5)


class Params(C.Structure):
    _fields_ = [("imp_locus", C.c_int32), ("imp_missing", C.c_int32), ("imp_sample", C.c_int32),
                ("ignorefilt", C.c_int32), ("use_cov", C.c_int32), ("skip_aftest", C.c_int32),
                ("mincs", C.c_int64), ("maxmis", C.c_double), ("afmisp", C.c_double)]


LOCUS_DTYPE = np.dtype([("klass", "<i4"), ("used", "<i4"), ("eaidx", "<i4"), ("reserved", "<i4"),
                        ("ngt", "<i8"), ("nmiss", "<i8"), ("neff", "<i8"), ("imputed", "<f8")], align=True)
ROW_DTYPE = np.dtype([("gt_row", "<i4"), ("eaidx", "<i4"), ("beta", "<f8"), ("eaf", "<f8"),
                      ("ref_is_ea", "<i4"), ("kind", "<i4")], align=True)
assert LOCUS_DTYPE.itemsize == 48 and ROW_DTYPE.itemsize == 32

_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(ROOT, "oracle", "nimpress_oracle.c")
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
        L = C.CDLL(_SO)
        for fn, args in (("orc_dbinom", [C.c_int64, C.c_int64, C.c_double]),
                         ("orc_pbinom", [C.c_int64, C.c_int64, C.c_double]),
                         ("orc_binom_test", [C.c_int64, C.c_int64, C.c_double]),
                         ("orc_betai", [C.c_double, C.c_double, C.c_double])):
            getattr(L, fn).restype = C.c_double
            getattr(L, fn).argtypes = args
        L.orc_format_float.argtypes = [C.c_double, C.c_char_p, C.c_int]
        L.orc_compute_scores_files.restype = C.c_int
        L.orc_compute_scores_files.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(Params),
                                               C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.c_void_p, C.c_int64,
                                               C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_char_p, C.c_int64,
                                               C.c_char_p, C.c_int64]
        L.orc_score_matrix.restype = C.c_int
        L.orc_score_matrix.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int64, C.c_void_p,
                                       C.c_int64, C.POINTER(Params), C.c_double, C.c_int32, C.c_void_p,
                                       C.c_void_p, C.POINTER(C.c_int64)]
        L.orc_synth_fill.restype = None
        L.orc_synth_fill.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_uint64,
                                     C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def params(imp_locus="ps", imp_missing="homref", imp_sample="int_ps", maxmis=0.05, afmisp=0.001, mincs=100,
           ignorefilt=False, cov=False, skip_aftest=False):
    return Params(LOCUS[imp_locus], MISSING[imp_missing], SAMPLE[imp_sample], int(ignorefilt), int(cov),
                  int(skip_aftest), int(mincs), float(maxmis), float(afmisp))


def abs_floor(betas, nloci):
    """Absolute part of the score tolerance for paths that add the reference's rounded products in another
    association: 8 * eps * sum|beta| / nloci.  Every term of a sample's sum is at most 2*|beta| in magnitude
    and the sum is divided by 2*nloci, so this is eight roundings of the largest value a normalised partial
    sum can take -- the scale of the arithmetic, not a constant."""
    b = np.asarray(betas, dtype=np.float64)
    b = b[np.isfinite(b)]
    return 8.0 * np.finfo(np.float64).eps * float(np.abs(b).sum()) / max(int(nloci), 1)


def _score_file_betas(path):
    out = []
    for ln in open(path).read().split("\n")[5:]:
        f = ln.rstrip().split("\t")
        if len(f) >= 5:
            try:
                out.append(float(f[4]))
            except ValueError:
                pass
    return out


def format_float(v):
    b = C.create_string_buffer(64)
    lib().orc_format_float(float(v), b, 64)
    return b.value.decode()


def compute_scores_files(score, vcf, bed=None, cap_samples=1 << 20, cap_loci=1 << 20, **kw):
    """computePolygenicScores on files -> dict(scores, loci, nloci, warn, samples)."""
    p = params(cov=bed is not None, **kw)
    scores = np.zeros(cap_samples, dtype=np.float64)
    loci = np.zeros(cap_loci, dtype=LOCUS_DTYPE)
    n, nl, used = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    warn = C.create_string_buffer(1 << 22)
    names = C.create_string_buffer(1 << 24)
    rc = lib().orc_compute_scores_files(os.fsencode(score), os.fsencode(vcf), os.fsencode(bed) if bed else None,
                                        C.byref(p), scores.ctypes.data, cap_samples, C.byref(n), loci.ctypes.data,
                                        cap_loci, C.byref(nl), C.byref(used), warn, len(warn), names, len(names))
    if rc:
        raise RuntimeError(f"oracle rc={rc}")
    return dict(scores=scores[:n.value].copy(), loci=loci[:nl.value].copy(), nloci=used.value,
                warn=warn.value.decode(), samples=names.value.decode().split("\n")[:n.value],
                abs_floor=abs_floor(_score_file_betas(score), used.value))


def score_matrix(gt, n_samples, ploidy, rows, offset=0.0, threads=1, **kw):
    """In-memory path.  gt: 2-D integer array [n_gt_rows, >= n_samples*ploidy] of dtype int8/16/32
    (C-contiguous rows).  rows: ROW_DTYPE array in processing order."""
    gt = np.asarray(gt)
    if gt.dtype == np.float32:                          # FORMAT/DS rows (not in the reference: parity unpinned)
        return score_matrix_ds(gt, n_samples, rows, offset, threads, **kw)
    assert gt.ndim == 2 and gt.dtype in (np.int8, np.int16, np.int32) and gt.strides[1] == gt.itemsize
    rows = np.ascontiguousarray(rows, dtype=ROW_DTYPE)
    p = params(**kw)
    scores = np.zeros(n_samples, dtype=np.float64)
    loci = np.zeros(len(rows), dtype=LOCUS_DTYPE)
    used = C.c_int64(0)
    rc = lib().orc_score_matrix(gt.ctypes.data, gt.itemsize, n_samples, ploidy, gt.strides[0], rows.ctypes.data,
                                len(rows), C.byref(p), float(offset), threads, scores.ctypes.data,
                                loci.ctypes.data, C.byref(used))
    if rc:
        raise RuntimeError(f"oracle rc={rc}")
    return dict(scores=scores, loci=loci, nloci=used.value, abs_floor=abs_floor(rows["beta"], used.value))


def score_matrix_ds(ds, n_samples, rows, offset=0.0, threads=1, **kw):
    """FORMAT/DS rows: ds float32 [n_rows, >= n_samples], one expected ALT dosage per sample (BCF float
    missing / NaN = no call).  loci["neff"] holds the bits of the fp64 dosage sum."""
    ds = np.asarray(ds)
    assert ds.ndim == 2 and ds.dtype == np.float32 and ds.strides[1] == 4
    rows = np.ascontiguousarray(rows, dtype=ROW_DTYPE)
    p = params(**kw)
    scores = np.zeros(n_samples, dtype=np.float64)
    loci = np.zeros(len(rows), dtype=LOCUS_DTYPE)
    used = C.c_int64(0)
    rc = lib().orc_score_matrix(ds.ctypes.data, -4, n_samples, 1, ds.strides[0], rows.ctypes.data, len(rows), C.byref(p),
                                float(offset), threads, scores.ctypes.data, loci.ctypes.data, C.byref(used))
    if rc:
        raise RuntimeError(f"oracle rc={rc}")
    return dict(scores=scores, loci=loci, nloci=used.value, abs_floor=abs_floor(rows["beta"], used.value))


def synth_fill(gt, n_samples, v0, seed, af_thr16, miss_thr24, alt_code):
    """Fill int8 rows gt[r, :2*n_samples] with the synthetic cohort for variants v0..v0+R."""
    assert gt.dtype == np.int8 and gt.ndim == 2 and gt.strides[1] == 1
    R = gt.shape[0]
    a = np.ascontiguousarray(af_thr16, dtype=np.uint32)
    m = np.ascontiguousarray(miss_thr24, dtype=np.uint32)
    c = np.ascontiguousarray(alt_code, dtype=np.int32)
    assert len(a) == len(m) == len(c) == R
    lib().orc_synth_fill(gt.ctypes.data, n_samples, gt.strides[0], v0, R, seed, a.ctypes.data, m.ctypes.data,
                         c.ctypes.data)
    return gt
