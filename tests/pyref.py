"""Tests-only: a second, independent restatement of the reference's scoring loop in plain Python
floats, written in the reference's own shape (a dosage list per locus with NaN for missing calls,
tallied and imputed in place) -- src/nimpress.nim:32-47 (tallyAlleles), :384-391 (getRawDosages),
:417-481 (imputeLocusDosages / imputeSampleDosages), :526-583 (getImputedDosages), :626-649
(computePolygenicScores).  It cross-checks the C oracle on random small inputs beyond the 13 golden
vectors (tests/test_oracle_crosscheck.py); far too slow for anything else."""
import math

NAN = float("nan")
KIND_GT, KIND_NOTCOV, KIND_ABSENT, KIND_FILTER, CLASS_MAXMIS = 0, 1, 2, 3, 4
_SENT = {1: (-128, -127), 2: (-32768, -32767), 4: (-2147483648, -2147483647)}   # (missing, vector_end) per width


def raw_dosages(values, ploidy, width, eaidx):
    """:384-391 on the stored integers: d = 0; allele == eaidx -> d += 1; allele == -1 -> NaN."""
    miss_s, vend = _SENT[width]
    out = []
    for s in range(len(values) // ploidy):
        d = 0.0
        for k in range(ploidy):
            raw = int(values[s * ploidy + k])
            if raw == vend:
                break
            if raw == miss_s:
                continue
            val = raw if raw < 0 else (raw >> 1) - 1
            if val == eaidx:
                d += 1.0
            elif val == -1:
                d = NAN
        out.append(d)
    return out


def tally(dosages):
    ng = nm = ne = 0.0
    for d in dosages:
        if math.isnan(d):
            nm += 1.0
        else:
            ng += 1.0
            ne += d
    return ng, nm, ne


def impute_locus(n, row, imp_locus):
    if imp_locus == "ignore":
        return None
    v = {"ps": row["eaf"] * 2.0, "homref": 2.0 if row["ref_is_ea"] else 0.0, "fail": NAN}[imp_locus]
    return [v] * n, v


def score(gt, n, ploidy, rows, offset, imp_locus="ps", imp_missing="homref", imp_sample="int_ps", maxmis=0.05, mincs=100):
    """gt: integer array [n_gt_rows, >= n*ploidy]; rows: sequence of dicts / structured records with
    gt_row, eaidx, beta, eaf, ref_is_ea, kind.  Returns (scores, nloci, loci) with loci as tuples
    (klass, used, ngt, nmiss, neff, imputed) -- ngt/nmiss/neff -1 when no tally was made."""
    width = gt.dtype.itemsize
    scores = [0.0] * n
    nloci = 0
    loci = []
    for row in rows:
        kind, tallies, res = int(row["kind"]), (-1, -1, -1), None
        if kind == KIND_GT and int(row["gt_row"]) < 0:
            kind = KIND_ABSENT
        if kind in (KIND_NOTCOV, KIND_FILTER):
            res = impute_locus(n, row, imp_locus)
            klass = kind
        elif kind == KIND_ABSENT:
            klass = kind
            if imp_missing == "homref":
                v = 2.0 if row["ref_is_ea"] else 0.0
                res = ([v] * n, v)
        else:
            dos = raw_dosages(gt[int(row["gt_row"])][:n * ploidy], ploidy, width, int(row["eaidx"]))
            ng, nm, ne = tally(dos)
            tallies = (int(ng), int(nm), int(ne))
            if nm / float(n) > maxmis:
                klass = CLASS_MAXMIS
                res = impute_locus(n, row, imp_locus)
            else:
                klass = KIND_GT
                if imp_sample in ("int_ps", "int_fail"):
                    v = ne / ng if ng >= float(mincs) else (row["eaf"] * 2.0 if imp_sample == "int_ps" else NAN)
                else:
                    v = {"ps": row["eaf"] * 2.0, "homref": 2.0 if row["ref_is_ea"] else 0.0, "fail": NAN}[imp_sample]
                res = ([v if math.isnan(d) else d for d in dos], v)
        if res is None:
            loci.append((klass, 0) + tallies + (NAN,))
            continue
        dosages, v = res
        beta = float(row["beta"])
        for i in range(n):
            scores[i] += dosages[i] * beta
        nloci += 1
        loci.append((klass, 1) + tallies + (v,))
    denom = float(nloci) * 2.0
    out = []
    for s in scores:
        q = s / denom if denom != 0.0 else (NAN if s == 0.0 or math.isnan(s) else math.copysign(math.inf, s))
        out.append(q + offset)
    return out, nloci, loci
