"""tests/test_set1.nim of the reference, line for line, against this engine: the same calls to
computePolygenicScores on the same fixtures with the same expected vectors and the same
checkFloats rule (NaN pattern equal, 1e-4 absolute).  Each case also runs in exact-order mode."""
import math
import os

import pytest

pytestmark = pytest.mark.gpu

S1 = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "set1")
NaN = math.nan


def checkFloats(x, target):                       # tests/test_set1.nim:14-22
    if len(x) != len(target):
        return False
    for xi, ti in zip(x, target):
        if math.isnan(ti) != math.isnan(xi):
            return False
        if not math.isnan(ti) and abs(ti - xi) > 1e-4:
            return False
    return True


@pytest.fixture(scope="module")
def N():
    import __graft_entry__ as g
    g.build()
    from nimpress_b200 import api
    return api


@pytest.fixture(params=[False, True], ids=["tile4", "exact-order"])
def setup(N, request):                            # suite "set1" setup (:26-33)
    genotypeVcf, scoreFile, coveredBed = N.VCF(), N.ScoreFile(), N.GenomeIntervals()
    assert N.open_vcf(genotypeVcf, os.path.join(S1, "set1.vcf.gz"))
    assert N.open_score(scoreFile, os.path.join(S1, "set1.score"))
    assert N.loadBedIntervals(coveredBed, os.path.join(S1, "set1.bed"))
    return N, [], scoreFile, genotypeVcf, coveredBed, request.param


L = lambda N: (N.ImputeMethodLocus, N.ImputeMethodMissing, N.ImputeMethodSample)


def test_locus_ps_mmr1_sample_fail(setup):        # :36-45
    N, scores, scoreFile, genotypeVcf, coveredBed, ex = setup
    Lo, Mi, Sa = L(N)
    N.computePolygenicScores(scores, scoreFile, genotypeVcf, False, coveredBed, Lo.ps, Mi.homref, Sa.fail,
                             maxMissingRate=1.0, afMismatchPthresh=1.0, minGtForInternalImput=100, ignoreFilterField=False, exactOrder=ex)
    assert checkFloats(scores, [NaN, 0.108, NaN, NaN, NaN, NaN])


def test_locus_ps_mmr02_sample_fail(setup):       # :48-57
    N, scores, scoreFile, genotypeVcf, coveredBed, ex = setup
    Lo, Mi, Sa = L(N)
    N.computePolygenicScores(scores, scoreFile, genotypeVcf, False, coveredBed, Lo.ps, Mi.homref, Sa.fail,
                             maxMissingRate=0.2, afMismatchPthresh=1.0, minGtForInternalImput=100, ignoreFilterField=False, exactOrder=ex)
    assert checkFloats(scores, [0.075166667, 0.1085, NaN, NaN, NaN, -0.0165])


def test_locus_ps_sample_homref(setup):           # :60-69
    N, scores, scoreFile, genotypeVcf, coveredBed, ex = setup
    Lo, Mi, Sa = L(N)
    N.computePolygenicScores(scores, scoreFile, genotypeVcf, False, coveredBed, Lo.ps, Mi.homref, Sa.homref,
                             maxMissingRate=0.2, afMismatchPthresh=1.0, minGtForInternalImput=100, ignoreFilterField=False, exactOrder=ex)
    assert checkFloats(scores, [0.075166667, 0.1085, 0.075166667, 0.141833333, 0.000166667, -0.0165])


def test_locus_ps_sample_int3_ps(setup):          # :72-81
    N, scores, scoreFile, genotypeVcf, coveredBed, ex = setup
    Lo, Mi, Sa = L(N)
    N.computePolygenicScores(scores, scoreFile, genotypeVcf, False, coveredBed, Lo.ps, Mi.homref, Sa.int_ps,
                             maxMissingRate=1.0, afMismatchPthresh=1.0, minGtForInternalImput=3, ignoreFilterField=False, exactOrder=ex)
    assert checkFloats(scores, [0.075166667, 0.108, 0.070166667, 0.036833333, 0.006833333, -0.0165])


def test_locus_ps_sample_int100_ps(setup):        # :84-93
    N, scores, scoreFile, genotypeVcf, coveredBed, ex = setup
    Lo, Mi, Sa = L(N)
    N.computePolygenicScores(scores, scoreFile, genotypeVcf, False, coveredBed, Lo.ps, Mi.homref, Sa.int_ps,
                             maxMissingRate=1.0, afMismatchPthresh=1.0, minGtForInternalImput=100, ignoreFilterField=False, exactOrder=ex)
    assert checkFloats(scores, [0.075166667, 0.108, 0.074333333, 0.140333333, 0.006833333, -0.0165])


def test_locus_ps_sample_int100_fail(setup):      # :96-105
    N, scores, scoreFile, genotypeVcf, coveredBed, ex = setup
    Lo, Mi, Sa = L(N)
    N.computePolygenicScores(scores, scoreFile, genotypeVcf, False, coveredBed, Lo.ps, Mi.homref, Sa.int_fail,
                             maxMissingRate=1.0, afMismatchPthresh=1.0, minGtForInternalImput=100, ignoreFilterField=False, exactOrder=ex)
    assert checkFloats(scores, [NaN, 0.108, NaN, NaN, NaN, NaN])


def test_locus_homref_mmr1_sample_fail(setup):    # :108-117
    N, scores, scoreFile, genotypeVcf, coveredBed, ex = setup
    Lo, Mi, Sa = L(N)
    N.computePolygenicScores(scores, scoreFile, genotypeVcf, False, coveredBed, Lo.homref, Mi.homref, Sa.fail,
                             maxMissingRate=1.0, afMismatchPthresh=1.0, minGtForInternalImput=100, ignoreFilterField=False, exactOrder=ex)
    assert checkFloats(scores, [NaN, 0.098, NaN, NaN, NaN, NaN])


def test_locus_homref_mmr02_sample_fail(setup):   # :120-129
    N, scores, scoreFile, genotypeVcf, coveredBed, ex = setup
    Lo, Mi, Sa = L(N)
    N.computePolygenicScores(scores, scoreFile, genotypeVcf, False, coveredBed, Lo.homref, Mi.homref, Sa.fail,
                             maxMissingRate=0.2, afMismatchPthresh=1.0, minGtForInternalImput=100, ignoreFilterField=False, exactOrder=ex)
    assert checkFloats(scores, [0.064666667, 0.098, NaN, NaN, NaN, -0.027])


def test_locus_homref_sample_homref(setup):       # :132-141
    N, scores, scoreFile, genotypeVcf, coveredBed, ex = setup
    Lo, Mi, Sa = L(N)
    N.computePolygenicScores(scores, scoreFile, genotypeVcf, False, coveredBed, Lo.homref, Mi.homref, Sa.homref,
                             maxMissingRate=1.0, afMismatchPthresh=1.0, minGtForInternalImput=100, ignoreFilterField=False, exactOrder=ex)
    assert checkFloats(scores, [0.064666667, 0.098, 0.064666667, 0.131333333, -0.010333333, -0.027])


def test_locus_fail_mmr1(setup):                  # :144-153
    N, scores, scoreFile, genotypeVcf, coveredBed, ex = setup
    Lo, Mi, Sa = L(N)
    N.computePolygenicScores(scores, scoreFile, genotypeVcf, False, coveredBed, Lo.fail, Mi.homref, Sa.fail,
                             maxMissingRate=1.0, afMismatchPthresh=1.0, minGtForInternalImput=100, ignoreFilterField=False, exactOrder=ex)
    assert checkFloats(scores, [NaN] * 6)


def test_locus_fail_mmr02(setup):                 # :156-165
    N, scores, scoreFile, genotypeVcf, coveredBed, ex = setup
    Lo, Mi, Sa = L(N)
    N.computePolygenicScores(scores, scoreFile, genotypeVcf, False, coveredBed, Lo.fail, Mi.homref, Sa.fail,
                             maxMissingRate=0.2, afMismatchPthresh=1.0, minGtForInternalImput=100, ignoreFilterField=False, exactOrder=ex)
    assert checkFloats(scores, [NaN] * 6)


def test_locus_ps_sample_ps_coverage_filtered(setup):   # :168-177
    N, scores, scoreFile, genotypeVcf, coveredBed, ex = setup
    Lo, Mi, Sa = L(N)
    N.computePolygenicScores(scores, scoreFile, genotypeVcf, True, coveredBed, Lo.ps, Mi.homref, Sa.ps,
                             maxMissingRate=1.0, afMismatchPthresh=1.0, minGtForInternalImput=100, ignoreFilterField=False, exactOrder=ex)
    assert checkFloats(scores, [0.081, 0.081, 0.081, 0.1545, 0.006, 0.006])
    assert genotypeVcf.samples == ["S1", "S2", "S3", "S4", "S5", "S6"]


def test_versus_plink_190_defaults(setup):        # :180-190
    N, scores, scoreFile, genotypeVcf, coveredBed, ex = setup
    Lo, Mi, Sa = L(N)
    N.computePolygenicScores(scores, scoreFile, genotypeVcf, False, coveredBed, Lo.ignore, Mi.ignore, Sa.int_ps,
                             maxMissingRate=1.0, afMismatchPthresh=1.0, minGtForInternalImput=0, ignoreFilterField=True, exactOrder=ex)
    # Offset of 0.123 is as PLINK does not use external offsets.
    assert checkFloats(scores, [0.123 - 0.03, 0.123 - 0.01, 0.123 - 0.076, 0.123 - 0.096, 0.123 - 0.132, 0.123 - 0.16])
    plink = [float(l.split()[5]) for l in open(os.path.join(S1, "set1.plink190.result")).read().splitlines()[1:]]
    assert checkFloats(scores, [0.123 + p for p in plink])    # tests/set1.plink190.result:2-7
