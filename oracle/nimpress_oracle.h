/*
 * nimpress_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the scoring path of mpinese/nimpress
 * (reference: src/nimpress.nim).  It is the parity checker for the CUDA path and the
 * "port" CPU baseline of bench.py.  Nothing in nimpress_b200/ (the product) may include,
 * link or call it; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs do.
 *
 * Parity pin: checked against every golden vector the reference's own tests hold for this
 * path -- the 13 expected score vectors of tests/test_set1.nim (incl. the PLINK 1.90 golden)
 * and the known answers of tests/test_stats.nim -- see tests/test_oracle_golden.py.
 * Behaviour that lives in third-party code absent from /root/reference (hts-nim >= 0.2.21,
 * htslib 1.10.2, lapper >= 0.1.5; nimpress.nimble:13-16, Dockerfile:32) is restated from the
 * published behaviour of those libraries and is pinned only as far as set1 exercises it
 * (SURVEY.md section 8c lists what stays unpinned: BCF input, int16/int32 GT, phased and
 * haploid calls, half-calls).
 */
#ifndef NIMPRESS_ORACLE_H
#define NIMPRESS_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* enums follow src/nimpress.nim:412-414 in declaration order */
enum { ORC_LOCUS_PS = 0, ORC_LOCUS_HOMREF = 1, ORC_LOCUS_FAIL = 2, ORC_LOCUS_IGNORE = 3 };
enum { ORC_MISSING_HOMREF = 0, ORC_MISSING_IGNORE = 1 };
enum { ORC_SAMPLE_PS = 0, ORC_SAMPLE_HOMREF = 1, ORC_SAMPLE_FAIL = 2, ORC_SAMPLE_INT_PS = 3,
       ORC_SAMPLE_INT_FAIL = 4 };

/* per-locus class: which branch of getImputedDosages (src/nimpress.nim:484-585) was taken */
enum { ORC_CLASS_OK = 0, ORC_CLASS_NOTCOV = 1, ORC_CLASS_ABSENT = 2, ORC_CLASS_FILTER = 3,
       ORC_CLASS_MAXMIS = 4 };

typedef struct {
    int32_t imp_locus;    /* --imp-locus   [ps]     */
    int32_t imp_missing;  /* --imp-missing [homref] */
    int32_t imp_sample;   /* --imp-sample  [int_ps] */
    int32_t ignorefilt;   /* --ignorefilt           */
    int32_t use_cov;      /* --cov given            */
    int32_t skip_aftest;  /* 1: skip the AF-mismatch binomTest (it only emits WARN lines) */
    int64_t mincs;        /* --mincs  [100]         */
    double  maxmis;       /* --maxmis [0.05]        */
    double  afmisp;       /* --afmisp [0.001]       */
} orc_params;

typedef struct {
    int32_t klass;        /* ORC_CLASS_*                                   */
    int32_t used;         /* 1 if the locus entered the sum / nloci        */
    int32_t eaidx;        /* effect-allele index in the record, -1 if none */
    int32_t reserved;
    int64_t ngt;          /* tallyAlleles results (exact integers); -1 when no tally was made */
    int64_t nmiss;
    int64_t neff;
    double  imputed;      /* OK: dosage given to missing samples; else the locus constant */
} orc_locus;

/* one score row of the in-memory entry point */
typedef struct {
    int32_t gt_row;       /* row of the genotype slab, -1 when the locus has no record   */
    int32_t eaidx;        /* a1: 0 = REF, k = k-th ALT                                    */
    double  beta;
    double  eaf;
    int32_t ref_is_ea;    /* refseq == easeq                                              */
    int32_t kind;         /* 0 record present and FILTER-passing; ORC_CLASS_NOTCOV /
                             ORC_CLASS_ABSENT / ORC_CLASS_FILTER decided by the caller    */
} orc_row;

/* ---- stats (src/nimpress.nim:51-188) ---- */
double orc_dbinom(int64_t x, int64_t n, double p);
double orc_betai(double a, double b, double x);
double orc_pbinom(int64_t x, int64_t n, double p);
double orc_binom_test(int64_t x, int64_t n, double p);

/* Nim's `$`(float) for Nim < 1.6: "%.16g", ".0" appended when integral, nan/inf/-inf. */
int orc_format_float(double v, char *buf, int buflen);

/*
 * Whole path on files (computePolygenicScores, src/nimpress.nim:592-649, plus the open/parse
 * code it depends on).  vcf_path: VCF text, plain or gzip/BGZF.  bed_path may be NULL.
 * scores_out: capacity cap_samples.  loci_out: capacity cap_loci (one per score row, file
 * order).  warn_buf receives the "WARN ..." lines the reference logs (may be NULL).
 * Returns 0, or a negative code (-1 cannot open genotypes, -2 cannot open score file,
 * -3 malformed input = the reference's doAssert/exception paths, -4 capacity).
 */
int orc_compute_scores_files(const char *score_path, const char *vcf_path, const char *bed_path,
                             const orc_params *p, double *scores_out, int64_t cap_samples,
                             int64_t *n_samples_out, orc_locus *loci_out, int64_t cap_loci,
                             int64_t *n_loci_out, int64_t *nloci_used_out,
                             char *warn_buf, int64_t warn_cap,
                             char *names_buf, int64_t names_cap);

/*
 * Same path on an in-memory genotype slab in BCF FORMAT/GT encoding
 * ((allele+1)<<1|phased; width 1, 2 or 4 bytes; sample-major, `ploidy` values per sample;
 * missing/vector_end sentinels of that width), rows `row_stride` bytes apart.  Rows are
 * processed in the order given, exactly like the reference's loop over the score file.
 * n_threads > 1 splits the ROW list into contiguous ranges (one partial score vector per
 * thread, summed in range order) -- this is NOT the reference's summation order and is only
 * used as the multi-core CPU baseline.
 */
int orc_score_matrix(const void *gt, int32_t gt_width, int64_t n_samples, int32_t ploidy,
                     int64_t row_stride, const orc_row *rows, int64_t n_rows,
                     const orc_params *p, double offset, int32_t n_threads,
                     double *scores_out, orc_locus *loci_out, int64_t *nloci_used_out);

/*
 * Synthetic cohort (tests/bench only): genotype of (variant v, sample s) is a pure function
 * of (seed, v, s): x = splitmix64(seed + v*K1 + s*K2); allele_i = 16-bit field i of x below
 * af_thr16[v] (Binomial(2, af_v)); missing when bits 32..55 are below miss_thr24[v] (1 in 8 of
 * those a half-call); bit 56 = phased.  Written as int8 BCF GT pairs; arrays indexed by row.  The CUDA generator in the product library must
 * produce identical bytes (tests/test_synth.py).
 */
void orc_synth_fill(int8_t *gt, int64_t n_samples, int64_t row_stride, int64_t v0, int64_t n_rows,
                    uint64_t seed, const uint32_t *af_thr16, const uint32_t *miss_thr24,
                    const int32_t *alt_code);

#ifdef __cplusplus
}
#endif
#endif
