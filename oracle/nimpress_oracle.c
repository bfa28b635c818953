/*
 * nimpress_oracle.c -- TEST INFRASTRUCTURE ONLY (see nimpress_oracle.h).
 *
 * Plain-C restatement of the scoring path of mpinese/nimpress, function by function, with the
 * reference's pass structure and summation order.  Each function cites the lines of
 * /root/reference/src/nimpress.nim it follows.  Compile with -ffp-contract=off (the reference
 * binary is x86-64 baseline code without FMA: `scores[i] += dosages[i]*beta` is a rounded
 * multiply followed by a rounded add).
 */
#define _GNU_SOURCE
#include "nimpress_oracle.h"

#include <ctype.h>
#include <errno.h>
#include <math.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#define VEC_END32 (INT32_MIN + 1)
#define MISSING32 (INT32_MIN)

/* ------------------------------------------------------------------------------------------
 * stats -- src/nimpress.nim:50-188 (only feeds WARN lines; never changes a score)
 * ---------------------------------------------------------------------------------------- */

/* :51 */
static double lbinom(int64_t n, int64_t k) {
    return lgamma((double)n + 1.0) - lgamma((double)k + 1.0) - lgamma((double)(n - k) + 1.0);
}

/* :54-60 */
double orc_dbinom(int64_t x, int64_t n, double p) {
    if ((x == 0 && p == 0.0) || (x == n && p == 1.0)) return 1.0;
    return exp(lbinom(n, x) + (double)x * log(p) + (double)(n - x) * log(1.0 - p));
}

/* :63-117, NRC continued fraction */
static double betacf(double a, double b, double x) {
    const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    const double FPMIN = 1.0e-30, EPS = 3.0e-7;
    const int MAXIT = 100;
    double c = 1.0, d = 1.0 - qab * x / qap, result;
    if (fabs(d) < FPMIN) d = FPMIN;
    d = 1.0 / d;
    result = d;
    for (int m = 1; m <= MAXIT; m++) {
        double mf = (double)m;
        double aa1 = mf * (b - mf) * x / ((qam + 2 * mf) * (a + 2 * mf));
        d = 1.0 + aa1 * d;
        if (fabs(d) < FPMIN) d = FPMIN;
        c = 1.0 + aa1 / c;
        if (fabs(c) < FPMIN) c = FPMIN;
        d = 1.0 / d;
        result *= d * c;
        double aa2 = -(a + mf) * (qab + mf) * x / ((a + 2 * mf) * (qap + 2 * mf));
        d = 1.0 + aa2 * d;
        if (fabs(d) < FPMIN) d = FPMIN;
        c = 1.0 + aa2 / c;
        if (fabs(c) < FPMIN) c = FPMIN;
        d = 1.0 / d;
        double del = d * c;
        result *= del;
        if (fabs(del - 1.0) < EPS) return result;
    }
    return NAN;
}

/* :120-134 */
double orc_betai(double a, double b, double x) {
    if (!(x >= 0.0 && x <= 1.0)) return NAN; /* reference: doAssert */
    if (a == 0.0 || b == 0.0) return INFINITY;
    if (x == 0.0) return 0.0;
    if (x == 1.0) return 1.0;
    double bt = exp(lgamma(a + b) - lgamma(a) - lgamma(b) + a * log(x) + b * log(1.0 - x));
    if (x < (a + 1.0) / (a + b + 2.0)) return bt * betacf(a, b, x) / a;
    return 1.0 - bt * betacf(b, a, 1.0 - x) / b;
}

/* :138-152 */
double orc_pbinom(int64_t x, int64_t n, double p) {
    if (x < 0) return 0.0;
    if (x == n) return 1.0;
    return 1.0 - orc_betai((double)x + 1.0, (double)(n - x), p);
}

/* :155-188 */
double orc_binom_test(int64_t x, int64_t n, double p) {
    if (p == 0.0) return x == 0 ? 1.0 : 0.0;
    if (p == 1.0) return x == n ? 1.0 : 0.0;
    double probx = orc_dbinom(x, n, p);
    double expected = (double)n * p;
    if (fabs((double)x / expected - 1.0) < 1.0e-6) return 1.0;
    if ((double)x < expected) {
        int64_t y = 0;
        for (int64_t xi = (int64_t)ceil(expected); xi <= n; xi++)
            if (orc_dbinom(xi, n, p) <= probx * (1.0 + 1.0e-7)) y++;
        return orc_pbinom(x, n, p) + (1.0 - orc_pbinom(n - y, n, p));
    } else {
        int64_t y = 0;
        for (int64_t xi = 0; xi <= (int64_t)floor(expected); xi++)
            if (orc_dbinom(xi, n, p) <= probx * (1.0 + 1.0e-7)) y++;
        return orc_pbinom(y - 1, n, p) + (1.0 - orc_pbinom(x - 1, n, p));
    }
}

/* Nim (<1.6) `$`(float): C "%.16g"; nan/inf spelled lower case; ".0" appended when the text
 * has neither '.', 'e' nor a non-digit (system/strmantle.nim nimFloatToStr).  Used by
 * src/nimpress.nim:753 and inside the WARN texts. */
int orc_format_float(double v, char *buf, int buflen) {
    if (isnan(v)) return snprintf(buf, buflen, "nan");
    if (isinf(v)) return snprintf(buf, buflen, v > 0 ? "inf" : "-inf");
    int n = snprintf(buf, buflen, "%.16g", v);
    int has = 0;
    for (int i = 0; i < n; i++) {
        if (buf[i] == ',') buf[i] = '.';
        if (buf[i] == '.' || buf[i] == 'e' || buf[i] == 'E' || isalpha((unsigned char)buf[i])) has = 1;
    }
    if (!has && n + 2 < buflen) { buf[n++] = '.'; buf[n++] = '0'; buf[n] = 0; }
    return n;
}

/* ------------------------------------------------------------------------------------------
 * small utilities
 * ---------------------------------------------------------------------------------------- */

typedef struct { char *p; int64_t len, cap; int overflow; } sbuf;

static void sb_printf(sbuf *b, const char *fmt, ...) {
    if (!b || !b->p) return;
    va_list ap;
    va_start(ap, fmt);
    int64_t room = b->cap - b->len;
    int n = vsnprintf(b->p + b->len, room > 0 ? (size_t)room : 0, fmt, ap);
    va_end(ap);
    if (n < 0 || n >= room) { b->overflow = 1; if (room > 0) b->p[b->len] = 0; return; }
    b->len += n;
}

/* Nim strip(leading = false): trailing chars of {' ', \t, \v, \r, \n, \f} removed */
static void rstrip(char *s) {
    size_t n = strlen(s);
    while (n && (s[n - 1] == ' ' || s[n - 1] == '\t' || s[n - 1] == '\v' || s[n - 1] == '\r' ||
                 s[n - 1] == '\n' || s[n - 1] == '\f'))
        s[--n] = 0;
}

/* split on a single char like Nim split('\t'): keeps empty fields. returns count */
static int split_char(char *s, char sep, char **out, int max) {
    int n = 0;
    out[n++] = s;
    for (char *c = s; *c; c++)
        if (*c == sep) {
            *c = 0;
            if (n < max) out[n] = c + 1;
            n++;
        }
    return n;
}

/* Nim parseFloat: whole string must be a float; nan / inf spellings accepted */
static int parse_float(const char *s, double *out) {
    if (!*s) return -1;
    char *end;
    errno = 0;
    double v = strtod(s, &end);
    if (end == s || *end) return -1;
    *out = v;
    return 0;
}

static int parse_int(const char *s, int64_t *out) {
    if (!*s) return -1;
    char *end;
    errno = 0;
    long long v = strtoll(s, &end, 10);
    if (end == s || *end) return -1;
    *out = v;
    return 0;
}

/* read one line of any length from a FILE (Nim readLine: strips \n and \r\n). NULL at EOF */
static char *read_line(FILE *f, char **buf, size_t *cap) {
    size_t len = 0;
    int c, any = 0;
    while ((c = fgetc(f)) != EOF) {
        any = 1;
        if (c == '\n') break;
        if (len + 2 > *cap) { *cap = *cap ? *cap * 2 : 256; *buf = (char *)realloc(*buf, *cap); }
        (*buf)[len++] = (char)c;
    }
    if (!any) return NULL;
    if (len + 1 > *cap) { *cap = len + 16; *buf = (char *)realloc(*buf, *cap); }
    if (len && (*buf)[len - 1] == '\r') len--;
    (*buf)[len] = 0;
    return *buf;
}

/* ------------------------------------------------------------------------------------------
 * score file -- src/nimpress.nim:195-254
 * ---------------------------------------------------------------------------------------- */

typedef struct {
    char *contig, *refseq, *easeq;
    int64_t pos;
    double beta, eaf;
} score_entry;

typedef struct {
    double offset;
    score_entry *e;
    int64_t n;
} score_file;

static int64_t entry_stop(const score_entry *s) { return s->pos + (int64_t)strlen(s->refseq) - 1; } /* :230-231 */

static void score_free(score_file *sf) {
    for (int64_t i = 0; i < sf->n; i++) { free(sf->e[i].contig); free(sf->e[i].refseq); free(sf->e[i].easeq); }
    free(sf->e);
    sf->e = NULL; sf->n = 0;
}

/* open (:233-244) + items (:247-254), read eagerly.  0 ok, -2 cannot open, -3 malformed */
static int score_load(const char *path, score_file *sf) {
    memset(sf, 0, sizeof(*sf));
    FILE *f = fopen(path, "rb");
    if (!f) return -2;
    char *buf = NULL; size_t cap = 0;
    int rc = 0;
    for (int h = 0; h < 5; h++) {          /* name, desc, cite, genomever, offset */
        if (!read_line(f, &buf, &cap)) { rc = -3; goto done; }
        if (h == 4) { rstrip(buf); if (parse_float(buf, &sf->offset)) { rc = -3; goto done; } }
    }
    int64_t capn = 0;
    while (read_line(f, &buf, &cap)) {
        rstrip(buf);
        char *parts[8];
        int np = split_char(buf, '\t', parts, 8);
        if (np != 6) { rc = -3; goto done; }                /* doAssert lineparts.len == 6 (:252) */
        if (sf->n == capn) { capn = capn ? capn * 2 : 64; sf->e = (score_entry *)realloc(sf->e, capn * sizeof(score_entry)); }
        score_entry *e = &sf->e[sf->n];
        memset(e, 0, sizeof(*e));
        if (parse_int(parts[1], &e->pos) || parse_float(parts[4], &e->beta) || parse_float(parts[5], &e->eaf)) { rc = -3; goto done; }
        e->contig = strdup(parts[0]); e->refseq = strdup(parts[2]); e->easeq = strdup(parts[3]);
        sf->n++;
    }
done:
    free(buf);
    fclose(f);
    if (rc) score_free(sf);
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * coverage BED -- src/nimpress.nim:262-345
 * ---------------------------------------------------------------------------------------- */

typedef struct { char *contig; int64_t start, stop; } bed_ival;
typedef struct { bed_ival *v; int64_t n; int init; } bed_set;

static void bed_free(bed_set *b) {
    for (int64_t i = 0; i < b->n; i++) free(b->v[i].contig);
    free(b->v);
    memset(b, 0, sizeof(*b));
}

/* loadBedIntervals (:278-308): every line, >= 3 tab fields, no header/comment handling */
static int bed_load(const char *path, bed_set *b) {
    memset(b, 0, sizeof(*b));
    FILE *f = fopen(path, "rb");
    if (!f) return -2;
    char *buf = NULL; size_t cap = 0;
    int64_t capn = 0;
    int rc = 0;
    while (read_line(f, &buf, &cap)) {
        rstrip(buf);
        char *parts[4];
        int np = split_char(buf, '\t', parts, 4);
        if (np < 3) { rc = -3; break; }                     /* doAssert lineparts.len >= 3 (:293) */
        if (b->n == capn) { capn = capn ? capn * 2 : 64; b->v = (bed_ival *)realloc(b->v, capn * sizeof(bed_ival)); }
        bed_ival *iv = &b->v[b->n];
        if (parse_int(parts[1], &iv->start) || parse_int(parts[2], &iv->stop)) { rc = -3; break; }
        iv->contig = strdup(parts[0]);
        b->n++;
    }
    free(buf);
    fclose(f);
    if (rc) { bed_free(b); return rc; }
    b->init = 1;
    return 0;
}

/* isVariantCovered (:313-345).  The lapper query find(pos-1, stop+1) (:337) only pre-filters:
 * an interval that `contains` (:310-311: start < pos and stop >= entry.stop) always overlaps
 * [pos-1, stop+1), so the containment test over all intervals of the contig is the answer. */
static int bed_covered(const bed_set *b, const score_entry *e, sbuf *warn) {
    int have_contig = 0;
    for (int64_t i = 0; i < b->n; i++)
        if (!strcmp(b->v[i].contig, e->contig)) { have_contig = 1; break; }
    if (!have_contig) {
        sb_printf(warn, "WARN Contig %s not present within the coverage BED file.\n", e->contig); /* :326 */
        return 0;
    }
    int64_t stop = entry_stop(e);
    for (int64_t i = 0; i < b->n; i++)
        if (!strcmp(b->v[i].contig, e->contig) && b->v[i].start < e->pos && b->v[i].stop >= stop) return 1;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * VCF text (plain / gzip / BGZF) -- stands in for hts-nim `open`, `query`, `REF`, `ALT`,
 * `FILTER`, `format.genotypes` (third-party, not under /root/reference)
 * ---------------------------------------------------------------------------------------- */

typedef struct {
    char *contig;
    int64_t pos;      /* 1-based */
    int64_t end;      /* 1-based inclusive: pos + rlen - 1, rlen = len(REF) or INFO/END */
    char *ref;
    char **alt; int n_alt;
    char *filter;     /* "." / "PASS" / names joined by ';' */
    int32_t *gt;      /* htslib-widened int32 [n_samples * ploidy], NULL if no GT */
    int ploidy;
} vcf_rec;

typedef struct {
    char **samples; int64_t n_samples;
    vcf_rec *r; int64_t n;
} vcf_file;

static void vcf_free(vcf_file *v) {
    for (int64_t i = 0; i < v->n_samples; i++) free(v->samples[i]);
    free(v->samples);
    for (int64_t i = 0; i < v->n; i++) {
        vcf_rec *r = &v->r[i];
        free(r->contig); free(r->ref); free(r->filter); free(r->gt);
        for (int k = 0; k < r->n_alt; k++) free(r->alt[k]);
        free(r->alt);
    }
    free(v->r);
    memset(v, 0, sizeof(*v));
}

static char *gz_read_line(gzFile g, char **buf, size_t *cap) {
    size_t len = 0;
    if (*cap < 65536) { *cap = 65536; *buf = (char *)realloc(*buf, *cap); }
    for (;;) {
        if (!gzgets(g, *buf + len, (int)(*cap - len))) { if (!len) return NULL; break; }
        len += strlen(*buf + len);
        if (len && (*buf)[len - 1] == '\n') break;
        if (len + 1 >= *cap) { *cap *= 2; *buf = (char *)realloc(*buf, *cap); }
        else if (gzeof(g)) break;
    }
    while (len && ((*buf)[len - 1] == '\n' || (*buf)[len - 1] == '\r')) len--;   /* set1.vcf.gz is CRLF */
    (*buf)[len] = 0;
    return *buf;
}

/* htslib vcf_parse_format, GT branch: alleles split on '/' '|'; '.' -> 0|phase; k -> (k+1)<<1|phase;
 * an empty field is one missing allele. Returns the number of alleles written (<= max). */
static int parse_gt_string(const char *t, int32_t *x, int max) {
    int l = 0, is_phased = 0;
    for (;;) {
        if (*t == '.') { ++t; if (l < max) x[l] = is_phased; l++; }
        else if (isdigit((unsigned char)*t)) {
            uint32_t val = 0;
            while (isdigit((unsigned char)*t)) { val = val * 10 + (uint32_t)(*t - '0'); ++t; }
            if (l < max) x[l] = (int32_t)(((val + 1) << 1) | (uint32_t)is_phased);
            l++;
        } else break;
        is_phased = (*t == '|');
        if (*t != '|' && *t != '/') break;
        ++t;
    }
    if (!l) { if (max > 0) x[0] = 0; l = 1; }
    return l;
}

#define MAX_PLOIDY 16

static int vcf_load(const char *path, vcf_file *v) {
    memset(v, 0, sizeof(*v));
    gzFile g = gzopen(path, "rb");
    if (!g) return -1;
    gzbuffer(g, 1 << 20);
    char *buf = NULL; size_t cap = 0;
    int rc = 0, have_header = 0;
    int64_t capn = 0;
    while (gz_read_line(g, &buf, &cap)) {
        if (buf[0] == '#') {
            if (buf[1] == '#') continue;
            /* #CHROM line: samples are columns 10+ */
            have_header = 1;
            int col = 0;
            char *c = buf;
            while (c) {
                char *nx = strchr(c, '\t');
                if (nx) *nx = 0;
                if (col >= 9) {
                    v->samples = (char **)realloc(v->samples, (v->n_samples + 1) * sizeof(char *));
                    v->samples[v->n_samples++] = strdup(c);
                }
                col++;
                c = nx ? nx + 1 : NULL;
            }
            continue;
        }
        if (!buf[0]) continue;
        if (!have_header) { rc = -1; break; }
        if (v->n == capn) { capn = capn ? capn * 2 : 64; v->r = (vcf_rec *)realloc(v->r, capn * sizeof(vcf_rec)); }
        vcf_rec *r = &v->r[v->n];
        memset(r, 0, sizeof(*r));
        /* fixed columns */
        char *f[9]; int nf = 0; char *c = buf, *rest = NULL;
        while (c && nf < 9) {
            char *nx = strchr(c, '\t');
            if (nx) *nx = 0;
            f[nf++] = c;
            c = nx ? nx + 1 : NULL;
        }
        rest = c;
        if (nf < 8) { rc = -3; break; }
        int64_t pos;
        if (parse_int(f[1], &pos)) { rc = -3; break; }
        r->contig = strdup(f[0]); r->pos = pos; r->ref = strdup(f[3]); r->filter = strdup(f[6]);
        int64_t rlen = (int64_t)strlen(f[3]);
        /* INFO/END overrides rlen (htslib vcf_parse + tbx readrec) */
        for (char *q = f[7]; q && *q;) {
            if (!strncmp(q, "END=", 4)) { int64_t e = atoll(q + 4); if (e >= pos) rlen = e - pos + 1; break; }
            q = strchr(q, ';'); if (q) q++;
        }
        r->end = pos + rlen - 1;
        if (strcmp(f[4], ".")) {
            char *a = f[4];
            while (a) {
                char *nx = strchr(a, ',');
                if (nx) *nx = 0;
                r->alt = (char **)realloc(r->alt, (r->n_alt + 1) * sizeof(char *));
                r->alt[r->n_alt++] = strdup(a);
                a = nx ? nx + 1 : NULL;
            }
        }
        v->n++;
        /* FORMAT: locate GT */
        int gt_idx = -1;
        if (nf == 9) {
            int k = 0; char *a = f[8];
            while (a) {
                char *nx = strchr(a, ':');
                if (nx) *nx = 0;
                if (!strcmp(a, "GT")) gt_idx = k;
                k++;
                a = nx ? nx + 1 : NULL;
            }
        }
        if (gt_idx < 0 || v->n_samples == 0) continue;       /* no GT: r->gt stays NULL */
        int32_t *tmp = (int32_t *)malloc((size_t)v->n_samples * MAX_PLOIDY * sizeof(int32_t));
        int *cnt = (int *)malloc((size_t)v->n_samples * sizeof(int));
        int maxp = 1;
        char *s = rest;
        for (int64_t i = 0; i < v->n_samples; i++) {
            const char *field = "";
            char *nx = NULL;
            if (s) {
                nx = strchr(s, '\t');
                if (nx) *nx = 0;
                /* pick sub-field gt_idx */
                char *a = s; int k = 0;
                while (a && k < gt_idx) { a = strchr(a, ':'); if (a) a++; k++; }
                if (a) { char *e = strchr(a, ':'); if (e) *e = 0; field = a; }
            }
            int l = parse_gt_string(field, tmp + i * MAX_PLOIDY, MAX_PLOIDY);
            if (l > MAX_PLOIDY) l = MAX_PLOIDY;
            cnt[i] = l;
            if (l > maxp) maxp = l;
            s = nx ? nx + 1 : NULL;
        }
        r->ploidy = maxp;
        r->gt = (int32_t *)malloc((size_t)v->n_samples * maxp * sizeof(int32_t));
        for (int64_t i = 0; i < v->n_samples; i++)
            for (int k = 0; k < maxp; k++)
                r->gt[i * maxp + k] = k < cnt[i] ? tmp[i * MAX_PLOIDY + k] : VEC_END32;
        free(tmp); free(cnt);
    }
    free(buf);
    gzclose(g);
    if (!have_header && !rc) rc = -1;
    if (rc) vcf_free(v);
    return rc;
}

/* findVariant (:353-364): records overlapping contig:pos-stop in file order; first with
 * REF == refseq and (easeq == refseq or easeq in ALT).  POS itself is never compared. */
static const vcf_rec *find_variant(const vcf_file *v, const score_entry *e) {
    int64_t stop = entry_stop(e);
    for (int64_t i = 0; i < v->n; i++) {
        const vcf_rec *r = &v->r[i];
        if (strcmp(r->contig, e->contig)) continue;
        if (r->pos > stop || r->end < e->pos) continue;
        if (!strcmp(r->ref, e->refseq)) {
            if (!strcmp(e->easeq, e->refseq)) return r;
            for (int k = 0; k < r->n_alt; k++)
                if (!strcmp(r->alt[k], e->easeq)) return r;
        }
    }
    return NULL;
}

/* ------------------------------------------------------------------------------------------
 * the hot path -- src/nimpress.nim:32-47, 367-391, 417-481, 484-585, 592-649
 * ---------------------------------------------------------------------------------------- */

/* hts-nim Allele.value: negative raw values (sentinels) are returned unchanged, else (a>>1)-1 */
static inline int64_t allele_value(int32_t a) { return a < 0 ? (int64_t)a : (int64_t)(a >> 1) - 1; }

/* getRawDosages (:367-391) on the widened int32 array */
static void get_raw_dosages(double *raw, const int32_t *gts, int64_t n, int ploidy, int eaidx) {
    for (int64_t i = 0; i < n; i++) {
        double d = 0.0;
        for (int k = 0; k < ploidy; k++) {
            int64_t val = allele_value(gts[i * ploidy + k]);
            if (val == eaidx) d += 1;
            else if (val == -1) d = NAN;
        }
        raw[i] = d;
    }
}

/* tallyAlleles (:32-47): three doubles holding exact integers */
static void tally_alleles(const double *raw, int64_t n, double *ngt, double *nmiss, double *neff) {
    double g = 0.0, m = 0.0, e = 0.0;
    for (int64_t i = 0; i < n; i++) {
        if (isnan(raw[i])) m += 1.0;
        else { g += 1.0; e += raw[i]; }
    }
    *ngt = g; *nmiss = m; *neff = e;
}

/* FORMAT/DS rows -- NOT IN THE REFERENCE (it reads GT only, :384; dosage input is "Future" in README.md:162-165):
 * parity unpinned.  Raw dosage = ds (effect allele = ALT) or 2 - ds (effect allele = REF); BCF float missing /
 * vector_end / NaN = no call.  neffectallele is a real-valued sum, so its ORDER is part of the definition
 * (nimpress_b200/csrc/npc_dosage.cuh): chunks of 8 samples left to right from +0.0; 32 consecutive chunk sums
 * combined by the butterfly v[i] += v[i ^ o], o = 16, 8, 4, 2, 1; the blocks of 256 samples left to right. */
static int ds_is_missing(uint32_t bits) {
    return bits == 0x7F800001u || bits == 0x7F800002u || (bits & 0x7FFFFFFFu) > 0x7F800000u;
}
static void get_raw_dosages_ds(double *raw, const float *ds, int64_t n, int eaidx) {
    for (int64_t i = 0; i < n; i++) {
        uint32_t bits;
        memcpy(&bits, &ds[i], 4);
        if (ds_is_missing(bits)) raw[i] = NAN;
        else raw[i] = eaidx == 0 ? 2.0 - (double)ds[i] : (double)ds[i];
    }
}
static void tally_dosages_ds(const double *raw, int64_t n, double *ngt, double *nmiss, double *neff) {
    double g = 0.0, m = 0.0, total = 0.0;
    for (int64_t b0 = 0; b0 < n; b0 += 256) {
        double v[32];
        for (int c = 0; c < 32; c++) {
            double s = 0.0;
            for (int k = 0; k < 8; k++) {
                int64_t i = b0 + c * 8 + k;
                if (i >= n) break;
                if (isnan(raw[i])) m += 1.0;
                else { g += 1.0; s = s + raw[i]; }
            }
            v[c] = s;
        }
        for (int o = 16; o; o >>= 1) {
            double w[32];
            for (int c = 0; c < 32; c++) w[c] = v[c] + v[c ^ o];
            memcpy(v, w, sizeof v);
        }
        total = total + v[0];
    }
    *ngt = g; *nmiss = m; *neff = total;
}

/* imputeLocusDosages (:417-447) */
static int impute_locus(double *dos, int64_t n, double eaf, int ref_is_ea, int method, double *value) {
    if (method == ORC_LOCUS_IGNORE) return 0;
    double v = NAN;
    if (method == ORC_LOCUS_PS) v = eaf * 2.0;
    else if (method == ORC_LOCUS_HOMREF) v = ref_is_ea ? 2.0 : 0.0;
    for (int64_t i = 0; i < n; i++) dos[i] = v;
    *value = v;
    return 1;
}

/* imputeSampleDosages (:450-481) */
static double impute_sample(double *dos, int64_t n, double eaf, int ref_is_ea, double neff, double ngt,
                            int64_t mincs, int method) {
    double v = NAN;
    switch (method) {
    case ORC_SAMPLE_PS: v = eaf * 2.0; break;
    case ORC_SAMPLE_HOMREF: v = ref_is_ea ? 2.0 : 0.0; break;
    case ORC_SAMPLE_FAIL: v = NAN; break;
    default:
        if (ngt >= (double)mincs) v = neff / ngt;
        else v = method == ORC_SAMPLE_INT_PS ? eaf * 2.0 : NAN;
    }
    for (int64_t i = 0; i < n; i++)
        if (isnan(dos[i])) dos[i] = v;
    return v;
}

typedef struct {
    const char *contig, *refseq, *easeq;   /* only for WARN texts (may be NULL) */
    int64_t pos, stop;
} locus_label;

/* the tail of getImputedDosages once the record (or its absence) is known: :536-585.
 * gts == NULL means "no record".  filter_fail: FILTER not in {".","PASS"} and !ignorefilt. */
static int locus_after_lookup_any(double *dos, int64_t n, const int32_t *gts, const float *ds, int ploidy, int eaidx,
                                  int filter_fail, const char *filter_str, double eaf, int ref_is_ea,
                                  const orc_params *p, const locus_label *lab, sbuf *warn, orc_locus *rec);
static int locus_after_lookup(double *dos, int64_t n, const int32_t *gts, int ploidy, int eaidx,
                              int filter_fail, const char *filter_str, double eaf, int ref_is_ea,
                              const orc_params *p, const locus_label *lab, sbuf *warn, orc_locus *rec) {
    return locus_after_lookup_any(dos, n, gts, NULL, ploidy, eaidx, filter_fail, filter_str, eaf, ref_is_ea, p, lab, warn, rec);
}
/* ds != NULL: a FORMAT/DS row (gts is then only a "record present" flag) */
static int locus_after_lookup_any(double *dos, int64_t n, const int32_t *gts, const float *ds, int ploidy, int eaidx,
                                  int filter_fail, const char *filter_str, double eaf, int ref_is_ea,
                                  const orc_params *p, const locus_label *lab, sbuf *warn, orc_locus *rec) {
    char fb[64], fb2[64];
    rec->eaidx = -1; rec->ngt = rec->nmiss = rec->neff = -1; rec->imputed = NAN;
    if (!gts) {                                                             /* :536-551 */
        rec->klass = ORC_CLASS_ABSENT;
        if (!p->skip_aftest && !isnan(eaf) && orc_binom_test(0, n * 2, eaf) < p->afmisp && lab->contig) {
            orc_format_float(eaf, fb, sizeof fb);
            sb_printf(warn, "WARN Variant %s:%lld:%s:%s cohort EAF is 0 in %lld samples.  This is highly"
                            " unlikely given polygenic score EAF of %s\n",
                      lab->contig, (long long)lab->pos, lab->refseq, lab->easeq, (long long)n, fb);
        }
        if (p->imp_missing == ORC_MISSING_HOMREF) {
            double v = ref_is_ea ? 2.0 : 0.0;
            for (int64_t i = 0; i < n; i++) dos[i] = v;
            rec->imputed = v;
            return 1;
        }
        return 0;
    }
    rec->eaidx = eaidx;
    if (filter_fail) {                                                      /* :553-558 */
        rec->klass = ORC_CLASS_FILTER;
        if (lab->contig)
            sb_printf(warn, "WARN Variant %s:%lld:%s:%s has a FILTER flag set (value \"%s\").  "
                            "Imputing all dosages at this locus.\n",
                      lab->contig, (long long)lab->pos, lab->refseq, lab->easeq, filter_str ? filter_str : "");
        return impute_locus(dos, n, eaf, ref_is_ea, p->imp_locus, &rec->imputed);
    }
    double ngt, nmiss, neff;
    if (ds) {
        get_raw_dosages_ds(dos, ds, n, eaidx);
        tally_dosages_ds(dos, n, &ngt, &nmiss, &neff);
        rec->ngt = (int64_t)ngt; rec->nmiss = (int64_t)nmiss; memcpy(&rec->neff, &neff, 8);   /* the fp64 sum's bits */
    } else {
        get_raw_dosages(dos, gts, n, ploidy, eaidx);                        /* :561 */
        tally_alleles(dos, n, &ngt, &nmiss, &neff);                         /* :563 */
        rec->ngt = (int64_t)ngt; rec->nmiss = (int64_t)nmiss; rec->neff = (int64_t)neff;
    }
    double missingrate = nmiss / (double)n;                                 /* :565 */
    if (missingrate > p->maxmis) {                                          /* :566-571 */
        rec->klass = ORC_CLASS_MAXMIS;
        if (lab->contig) {
            orc_format_float(missingrate * 100, fb, sizeof fb);
            sb_printf(warn, "WARN Locus %s:%lld-%lld has %s%% of samples missing a genotype. This exceeds "
                            "the missingness threshold; imputing all dosages at this locus.\n",
                      lab->contig, (long long)lab->pos, (long long)lab->stop, fb);
        }
        return impute_locus(dos, n, eaf, ref_is_ea, p->imp_locus, &rec->imputed);
    }
    if (!ds && !p->skip_aftest && !isnan(eaf) &&                            /* :573-579 */
        orc_binom_test((int64_t)neff, (n - (int64_t)nmiss) * 2, eaf) < p->afmisp && lab->contig) {
        orc_format_float(neff / (double)((n - (int64_t)nmiss) * 2), fb, sizeof fb);
        orc_format_float(eaf, fb2, sizeof fb2);
        sb_printf(warn, "WARN Variant %s:%lld:%s:%s cohort EAF is %s in %lld samples.  This is highly "
                        "unlikely given polygenic score EAF of %s\n",
                  lab->contig, (long long)lab->pos, lab->refseq, lab->easeq, fb, (long long)n, fb2);
    }
    rec->klass = ORC_CLASS_OK;
    rec->imputed = impute_sample(dos, n, eaf, ref_is_ea, neff, ngt, p->mincs, p->imp_sample); /* :582 */
    return 1;
}

/* computePolygenicScores epilogue (:643-649) */
static void normalise(double *scores, int64_t n, int64_t nloci, double offset) {
    for (int64_t i = 0; i < n; i++) scores[i] /= (double)nloci * 2.0;
    for (int64_t i = 0; i < n; i++) scores[i] += offset;
}

int orc_compute_scores_files(const char *score_path, const char *vcf_path, const char *bed_path,
                             const orc_params *p, double *scores_out, int64_t cap_samples,
                             int64_t *n_samples_out, orc_locus *loci_out, int64_t cap_loci,
                             int64_t *n_loci_out, int64_t *nloci_used_out,
                             char *warn_buf, int64_t warn_cap, char *names_buf, int64_t names_cap) {
    vcf_file vcf; score_file sf; bed_set bed;
    memset(&bed, 0, sizeof bed);
    int rc = vcf_load(vcf_path, &vcf);                       /* main :728 */
    if (rc) return rc;
    rc = score_load(score_path, &sf);                        /* main :732 */
    if (rc) { vcf_free(&vcf); return rc; }
    sbuf warn = { warn_buf, 0, warn_cap, 0 };
    if (warn_buf && warn_cap > 0) warn_buf[0] = 0;
    if (p->use_cov && bed_path) {
        int brc = bed_load(bed_path, &bed);                  /* main :737-740: failure is logged, not fatal */
        if (brc == -3) { vcf_free(&vcf); score_free(&sf); return -3; }
        if (brc) sb_printf(&warn, "FATAL Could not open coverage BED file %s\n", bed_path);
    }
    int64_t n = vcf.n_samples;
    if (n > cap_samples || sf.n > cap_loci) { vcf_free(&vcf); score_free(&sf); bed_free(&bed); return -4; }
    *n_samples_out = n; *n_loci_out = sf.n;
    if (names_buf) {
        sbuf nb = { names_buf, 0, names_cap, 0 };
        names_buf[0] = 0;
        for (int64_t i = 0; i < n; i++) sb_printf(&nb, "%s\n", vcf.samples[i]);
    }
    for (int64_t i = 0; i < n; i++) scores_out[i] = 0.0;     /* :626-628 */
    int64_t nloci = 0;
    double *dos = (double *)malloc((size_t)(n ? n : 1) * sizeof(double));
    for (int64_t li = 0; li < sf.n && !rc; li++) {           /* :634 */
        const score_entry *e = &sf.e[li];
        orc_locus *rec = &loci_out[li];
        memset(rec, 0, sizeof *rec);
        locus_label lab = { e->contig, e->refseq, e->easeq, e->pos, entry_stop(e) };
        int ref_is_ea = !strcmp(e->refseq, e->easeq);
        int used;
        if (p->use_cov && !bed_covered(&bed, e, &warn)) {    /* :526-531 */
            sb_printf(&warn, "WARN Locus %s:%lld-%lld is not covered by the sequence coverage BED.  "
                             "Imputing all dosages at this locus.\n", e->contig, (long long)e->pos, (long long)lab.stop);
            rec->klass = ORC_CLASS_NOTCOV; rec->eaidx = -1; rec->ngt = rec->nmiss = rec->neff = -1; rec->imputed = NAN;
            used = impute_locus(dos, n, e->eaf, ref_is_ea, p->imp_locus, &rec->imputed);
        } else {
            const vcf_rec *r = find_variant(&vcf, e);        /* :533 */
            int eaidx = -1, filter_fail = 0;
            if (r) {
                if (!r->gt) { rc = -3; break; }              /* reference: nil Genotypes -> crash */
                if (ref_is_ea) eaidx = 0;                    /* :375-379 */
                else for (int k = 0; k < r->n_alt; k++) if (!strcmp(r->alt[k], e->easeq)) { eaidx = k + 1; break; }
                filter_fail = !p->ignorefilt && strcmp(r->filter, ".") && strcmp(r->filter, "PASS");
            }
            used = locus_after_lookup(dos, n, r ? r->gt : NULL, r ? r->ploidy : 0, eaidx, filter_fail,
                                      r ? r->filter : NULL, e->eaf, ref_is_ea, p, &lab, &warn, rec);
        }
        rec->used = used;
        if (used) {                                          /* :639-641 */
            for (int64_t i = 0; i < n; i++) scores_out[i] += dos[i] * e->beta;
            nloci += 1;
        }
    }
    if (!rc) normalise(scores_out, n, nloci, sf.offset);
    if (nloci_used_out) *nloci_used_out = nloci;
    free(dos);
    vcf_free(&vcf); score_free(&sf); bed_free(&bed);
    return rc;
}

/* htslib bcf_get_format_values widening of one GT row to int32: width-specific missing /
 * vector_end sentinels become the int32 ones and everything after a vector_end in a sample is
 * vector_end.  This pass is part of the reference's per-locus cost. */
static void widen_row(int32_t *dst, const void *src, int width, int64_t n, int ploidy) {
    for (int64_t i = 0; i < n; i++) {
        int k = 0;
        for (; k < ploidy; k++) {
            int32_t a;
            int64_t j = i * ploidy + k;
            if (width == 1) { int8_t b = ((const int8_t *)src)[j]; a = b == INT8_MIN ? MISSING32 : b == INT8_MIN + 1 ? VEC_END32 : b; }
            else if (width == 2) { int16_t b = ((const int16_t *)src)[j]; a = b == INT16_MIN ? MISSING32 : b == INT16_MIN + 1 ? VEC_END32 : b; }
            else a = ((const int32_t *)src)[j];
            if (a == VEC_END32) break;
            dst[j] = a;
        }
        for (; k < ploidy; k++) dst[i * ploidy + k] = VEC_END32;
    }
}

typedef struct {
    const void *gt; int32_t gt_width; int64_t n; int32_t ploidy; int64_t row_stride;
    const orc_row *rows; int64_t r0, r1; const orc_params *p;
    double *scores; orc_locus *loci; int64_t nloci;
} mt_job;

static void *matrix_range(void *arg) {
    mt_job *j = (mt_job *)arg;
    int64_t n = j->n;
    double *dos = (double *)malloc((size_t)(n ? n : 1) * sizeof(double));
    int32_t *wide = (int32_t *)malloc((size_t)(n ? n : 1) * j->ploidy * sizeof(int32_t));
    locus_label lab = { NULL, NULL, NULL, 0, 0 };
    for (int64_t i = 0; i < n; i++) j->scores[i] = 0.0;
    for (int64_t ri = j->r0; ri < j->r1; ri++) {
        const orc_row *row = &j->rows[ri];
        orc_locus *rec = &j->loci[ri];
        memset(rec, 0, sizeof *rec);
        int used;
        if (row->kind == ORC_CLASS_NOTCOV) {
            rec->klass = ORC_CLASS_NOTCOV; rec->eaidx = -1; rec->ngt = rec->nmiss = rec->neff = -1; rec->imputed = NAN;
            used = impute_locus(dos, n, row->eaf, row->ref_is_ea, j->p->imp_locus, &rec->imputed);
        } else {
            const int32_t *g = NULL;
            if (row->kind == ORC_CLASS_FILTER) g = wide;      /* a FILTER-failed record is never decoded (:553-558) */
            else if (row->kind != ORC_CLASS_ABSENT && row->gt_row >= 0) {
                if (j->gt_width != -4) widen_row(wide, (const char *)j->gt + row->gt_row * j->row_stride, j->gt_width, n, j->ploidy);
                g = wide;
            }
            const float *dsrow = j->gt_width == -4 && g && row->kind != ORC_CLASS_FILTER
                                     ? (const float *)((const char *)j->gt + row->gt_row * j->row_stride) : NULL;
            used = locus_after_lookup_any(dos, n, g, dsrow, j->ploidy, row->eaidx, row->kind == ORC_CLASS_FILTER, NULL,
                                          row->eaf, row->ref_is_ea, j->p, &lab, NULL, rec);
        }
        rec->used = used;
        if (used) {
            double beta = row->beta;
            for (int64_t i = 0; i < n; i++) j->scores[i] += dos[i] * beta;
            j->nloci += 1;
        }
    }
    free(dos); free(wide);
    return NULL;
}

int orc_score_matrix(const void *gt, int32_t gt_width, int64_t n_samples, int32_t ploidy,
                     int64_t row_stride, const orc_row *rows, int64_t n_rows,
                     const orc_params *p, double offset, int32_t n_threads,
                     double *scores_out, orc_locus *loci_out, int64_t *nloci_used_out) {
    if (gt_width != 1 && gt_width != 2 && gt_width != 4 && gt_width != -4) return -3;   /* -4: fp32 FORMAT/DS rows (ploidy 1) */
    if (ploidy < 1 || (gt_width == -4 && ploidy != 1)) return -3;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_rows) n_threads = n_rows > 0 ? (int32_t)n_rows : 1;
    orc_params pp = *p;
    pp.skip_aftest = 1;                      /* no labels here: WARN lines are a file-level feature */
    mt_job *jobs = (mt_job *)calloc((size_t)n_threads, sizeof(mt_job));
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int t = 0; t < n_threads; t++) {
        mt_job *j = &jobs[t];
        j->gt = gt; j->gt_width = gt_width; j->n = n_samples; j->ploidy = ploidy; j->row_stride = row_stride;
        j->rows = rows; j->r0 = n_rows * t / n_threads; j->r1 = n_rows * (t + 1) / n_threads; j->p = &pp;
        j->loci = loci_out;
        j->scores = t == 0 ? scores_out : (double *)malloc((size_t)(n_samples ? n_samples : 1) * sizeof(double));
        if (n_threads > 1) pthread_create(&th[t], NULL, matrix_range, j);
        else matrix_range(j);
    }
    int64_t nloci = 0;
    for (int t = 0; t < n_threads; t++) {
        if (n_threads > 1) pthread_join(th[t], NULL);
        nloci += jobs[t].nloci;
        if (t > 0) {
            for (int64_t i = 0; i < n_samples; i++) scores_out[i] += jobs[t].scores[i];
            free(jobs[t].scores);
        }
    }
    normalise(scores_out, n_samples, nloci, offset);
    if (nloci_used_out) *nloci_used_out = nloci;
    free(jobs); free(th);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * synthetic cohort generator (tests / bench only; not reference behaviour)
 * ---------------------------------------------------------------------------------------- */

static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

void orc_synth_fill(int8_t *gt, int64_t n_samples, int64_t row_stride, int64_t v0, int64_t n_rows,
                    uint64_t seed, const uint32_t *af_thr16, const uint32_t *miss_thr24,
                    const int32_t *alt_code) {
    for (int64_t r = 0; r < n_rows; r++) {
        int8_t *row = gt + r * row_stride;
        uint64_t vkey = seed + (uint64_t)(v0 + r) * 0xD1B54A32D192ED03ULL;
        uint32_t aft = af_thr16[r], mt = miss_thr24[r];
        int8_t alt = (int8_t)((alt_code[r] + 1) << 1);
        for (int64_t s = 0; s < n_samples; s++) {
            uint64_t x = splitmix64(vkey + (uint64_t)s * 0x8CB92BA72F3D8DD7ULL);
            uint32_t u0 = (uint32_t)(x & 0xFFFF), u1 = (uint32_t)((x >> 16) & 0xFFFF);
            uint32_t um = (uint32_t)((x >> 32) & 0xFFFFFF);
            int8_t ph = (int8_t)((x >> 56) & 1);
            int8_t a0 = u0 < aft ? alt : 2, a1 = (int8_t)((u1 < aft ? alt : 2) | ph);
            if (um < mt) {
                a0 = 0;
                if (((x >> 57) & 7) != 0) a1 = ph;      /* 7 in 8: "./.", 1 in 8: half-call "./a" */
            }
            row[2 * s] = a0; row[2 * s + 1] = a1;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * optional CLI: nimpress_oracle [options] <scoredef> <genotypes.vcf[.gz]>   (main :652-753)
 * ---------------------------------------------------------------------------------------- */
#ifdef ORC_MAIN
static int parse_enum(const char *v, const char *const *names, int n) {
    for (int i = 0; i < n; i++) if (!strcmp(v, names[i])) return i;
    return -1;
}
int main(int argc, char **argv) {
    static const char *const L[] = { "ps", "homref", "fail", "ignore" };
    static const char *const M[] = { "homref", "ignore" };
    static const char *const S[] = { "ps", "homref", "fail", "int_ps", "int_fail" };
    orc_params p = { ORC_LOCUS_PS, ORC_MISSING_HOMREF, ORC_SAMPLE_INT_PS, 0, 0, 0, 100, 0.05, 0.001 };
    const char *cov = NULL, *pos[2] = { 0, 0 };
    int np = 0;
    for (int i = 1; i < argc; i++) {
        const char *a = argv[i];
        if (!strncmp(a, "--cov=", 6)) { cov = a + 6; p.use_cov = 1; }
        else if (!strncmp(a, "--imp-locus=", 12)) p.imp_locus = parse_enum(a + 12, L, 4);
        else if (!strncmp(a, "--imp-missing=", 14)) p.imp_missing = parse_enum(a + 14, M, 2);
        else if (!strncmp(a, "--imp-sample=", 13)) p.imp_sample = parse_enum(a + 13, S, 5);
        else if (!strncmp(a, "--maxmis=", 9)) p.maxmis = atof(a + 9);
        else if (!strncmp(a, "--mincs=", 8)) p.mincs = atoll(a + 8);
        else if (!strncmp(a, "--afmisp=", 9)) p.afmisp = atof(a + 9);
        else if (!strcmp(a, "--ignorefilt")) p.ignorefilt = 1;
        else if (!strcmp(a, "--no-aftest")) p.skip_aftest = 1;
        else if (np < 2) pos[np++] = a;
    }
    if (np != 2 || p.imp_locus < 0 || p.imp_missing < 0 || p.imp_sample < 0) {
        fprintf(stderr, "usage: nimpress_oracle [options] <scoredef> <genotypes.vcf>\n");
        return 1;
    }
    int64_t cap = 1 << 22, n = 0, nl = 0, used = 0;
    double *scores = (double *)malloc(cap * sizeof(double));
    orc_locus *loci = (orc_locus *)malloc((size_t)(1 << 22) * sizeof(orc_locus));
    char *warn = (char *)malloc(1 << 24), *names = (char *)malloc(1 << 26);
    int rc = orc_compute_scores_files(pos[0], pos[1], cov, &p, scores, cap, &n, loci, 1 << 22, &nl, &used,
                                      warn, 1 << 24, names, 1 << 26);
    if (rc == -1) { printf("FATAL Could not open input VCF file %s\n", pos[1]); return 255; }
    if (rc == -2) { printf("FATAL Could not open polygenic score file %s\n", pos[0]); return 255; }
    if (rc) return 1;
    fputs(warn, stdout);
    char *nm = names;
    for (int64_t i = 0; i < n; i++) {
        char *e = strchr(nm, '\n'); if (e) *e = 0;
        char fb[64]; orc_format_float(scores[i], fb, sizeof fb);
        printf("%s\t%s\n", nm, fb);
        nm = e ? e + 1 : nm;
    }
    return 0;
}
#endif
