// host_api.cpp -- C ABI of libnimpress_host.so (include/nimpress_host.h) and the command line.
#include <cstring>
#include <unistd.h>
#include <iostream>
#include <string>
#include <vector>

#include "../../include/nimpress_host.h"
#include "driver.hpp"
#include "fast_inflate.hpp"
#include "stats.hpp"

using namespace nph;

struct nph_result { ScoreResult r; };

static thread_local std::string g_err;

const char *nph_last_error(void) { return g_err.c_str(); }

static ScoreParams to_params(const nph_params *p) {
    ScoreParams q;
    q.imp_locus = p->imp_locus; q.imp_missing = p->imp_missing; q.imp_sample = p->imp_sample;
    q.ignorefilt = p->ignorefilt != 0; q.use_cov = p->use_cov != 0; q.device = p->device; q.exact_order = p->exact_order != 0;
    q.mincs = p->mincs; q.maxmis = p->maxmis; q.afmisp = p->afmisp;
    for (int d = 0; d < 32; d++) if (((uint32_t)p->device_mask >> d) & 1u) q.devices.push_back(d);
    q.use_ds = p->use_ds != 0;
    return q;
}

// open(ScoreFile) + loadBedIntervals as main() sequences them (:728-740)
static int load_inputs(const char *score_path, const char *bed_path, ScoreParams &q, ScoreFile &sf, GenomeIntervals &cov,
                       std::string *fatal) {
    if (!sf.load(score_path)) return NPH_EOPEN_SCORE;
    if (q.use_cov && bed_path) {
        if (!cov.load(bed_path) && fatal)            // the reference logs FATAL but carries on (:739-740)
            *fatal += std::string("FATAL Could not open coverage BED file ") + bed_path + "\n";
    }
    return NPH_OK;
}

int nph_compute_polygenic_scores(const char *score_path, const char *genotype_path, const char *bed_path,
                                 const nph_params *p, nph_result **out) {
    if (!score_path || !genotype_path || !p || !out) return NPH_EINPUT;
    *out = nullptr;
    try {
        ScoreParams q = to_params(p);
        {   // main() opens the VCF first (:728): an unreadable VCF wins over an unreadable score file
            std::unique_ptr<VariantSource> probe = open_variant_source(genotype_path);
            if (!probe) { g_err = std::string("Could not open input VCF file ") + genotype_path; return NPH_EOPEN_VCF; }
        }
        ScoreFile sf; GenomeIntervals cov;
        std::string fatal;
        int rc = load_inputs(score_path, bed_path, q, sf, cov, &fatal);
        if (rc) { g_err = std::string("Could not open polygenic score file ") + score_path; return rc; }
        nph_result *res = new nph_result();
        if (!compute_polygenic_scores(sf, genotype_path, cov, q, res->r)) {
            delete res;
            g_err = std::string("Could not open input VCF file ") + genotype_path;
            return NPH_EOPEN_VCF;
        }
        res->r.warnings = fatal + res->r.warnings;
        *out = res;
        return NPH_OK;
    } catch (const InputError &e) {
        g_err = e.what();
        return NPH_EINPUT;
    } catch (const std::exception &e) {
        g_err = e.what();
        return NPH_EGPU;
    }
}

int nph_compute_polygenic_scores_multi(const char *const *score_paths, int32_t n_scores, const char *genotype_path,
                                       const char *bed_path, const nph_params *p, nph_result **out) {
    if (!score_paths || n_scores < 1 || !genotype_path || !p || !out) return NPH_EINPUT;
    for (int32_t k = 0; k < n_scores; k++) out[k] = nullptr;
    try {
        ScoreParams q = to_params(p);
        {
            std::unique_ptr<VariantSource> probe = open_variant_source(genotype_path);
            if (!probe) { g_err = std::string("Could not open input VCF file ") + genotype_path; return NPH_EOPEN_VCF; }
        }
        std::vector<ScoreFile> files((size_t)n_scores);
        std::vector<const ScoreFile *> ptrs;
        GenomeIntervals cov;
        std::string fatal;
        for (int32_t k = 0; k < n_scores; k++) {
            if (!files[k].load(score_paths[k])) { g_err = std::string("Could not open polygenic score file ") + score_paths[k]; return NPH_EOPEN_SCORE; }
            ptrs.push_back(&files[k]);
        }
        if (q.use_cov && bed_path && !cov.load(bed_path)) fatal = std::string("FATAL Could not open coverage BED file ") + bed_path + "\n";
        std::vector<ScoreResult> rs;
        if (!compute_polygenic_scores_multi(ptrs, genotype_path, cov, q, rs)) {
            g_err = std::string("Could not open input VCF file ") + genotype_path;
            return NPH_EOPEN_VCF;
        }
        for (int32_t k = 0; k < n_scores; k++) {
            out[k] = new nph_result();
            out[k]->r = std::move(rs[k]);
            out[k]->r.warnings = fatal + out[k]->r.warnings;
        }
        return NPH_OK;
    } catch (const InputError &e) {
        g_err = e.what();
        return NPH_EINPUT;
    } catch (const std::exception &e) {
        g_err = e.what();
        return NPH_EGPU;
    }
}

int64_t nph_result_n_samples(const nph_result *r) { return (int64_t)r->r.scores.size(); }
int64_t nph_result_n_loci(const nph_result *r) { return (int64_t)r->r.loci.size(); }
int64_t nph_result_nloci_used(const nph_result *r) { return r->r.nloci; }
int64_t nph_result_rounds(const nph_result *r) { return r->r.rounds; }
int64_t nph_result_devices(const nph_result *r) { return r->r.devices; }
int64_t nph_result_records_read(const nph_result *r) { return r->r.records_read; }
int64_t nph_result_index_seeks(const nph_result *r) { return r->r.index_seeks; }
const double *nph_result_scores(const nph_result *r) { return r->r.scores.data(); }
const npc_locus *nph_result_loci(const nph_result *r) { return r->r.loci.data(); }
const char *nph_result_sample(const nph_result *r, int64_t i) { return r->r.samples[(size_t)i].c_str(); }
const char *nph_result_warnings(const nph_result *r) { return r->r.warnings.c_str(); }
void nph_result_free(nph_result *r) { delete r; }

int nph_plan(const char *score_path, const char *genotype_path, const char *bed_path, const nph_params *p,
             int32_t *kind_out, int32_t *eaidx_out, int64_t cap, int64_t *n_rows_out, int64_t *n_samples_out) {
    try {
        ScoreParams q = to_params(p);
        std::unique_ptr<VariantSource> vcf = open_variant_source(genotype_path);
        if (!vcf) return NPH_EOPEN_VCF;
        ScoreFile sf; GenomeIntervals cov;
        int rc = load_inputs(score_path, bed_path, q, sf, cov, nullptr);
        if (rc) return rc;
        if ((int64_t)sf.entries.size() > cap) return NPH_ECAPACITY;
        Matcher M(sf, cov, q);
        VariantRecord rec;
        while (vcf->next(rec)) M.match(rec);
        M.finish();
        for (size_t i = 0; i < sf.entries.size(); i++) { kind_out[i] = M.kind[i]; eaidx_out[i] = M.eaidx[i]; }
        *n_rows_out = (int64_t)sf.entries.size();
        *n_samples_out = vcf->n_samples();
        return NPH_OK;
    } catch (const InputError &e) {
        g_err = e.what();
        return NPH_EINPUT;
    }
}

int nph_plan_indexed(const char *score_path, const char *genotype_path, const char *bed_path, const nph_params *p,
                     int32_t *kind_out, int32_t *eaidx_out, int64_t cap, int64_t *n_rows_out, int64_t *n_samples_out,
                     int64_t *records_read_out, int64_t *index_seeks_out) {
    try {
        ScoreParams q = to_params(p);
        ScoreFile sf; GenomeIntervals cov;
        int rc = load_inputs(score_path, bed_path, q, sf, cov, nullptr);
        if (rc) return rc;
        if ((int64_t)sf.entries.size() > cap) return NPH_ECAPACITY;
        Matcher M(sf, cov, q);
        std::unique_ptr<VariantSource> vcf = open_variant_source(genotype_path, true);      // the driver's order: index first
        if (!vcf) return NPH_EOPEN_VCF;
        std::unordered_map<std::string, std::vector<std::pair<int64_t, int64_t>>> spans;
        M.add_spans(spans);
        if (!vcf->use_regions(spans, genotype_path)) {
            vcf = open_variant_source(genotype_path, false);
            if (!vcf) return NPH_EOPEN_VCF;
        }
        VariantRecord rec;
        int64_t nrec = 0;
        while (vcf->next(rec)) { nrec++; M.match(rec); }
        M.finish();
        for (size_t i = 0; i < sf.entries.size(); i++) { kind_out[i] = M.kind[i]; eaidx_out[i] = M.eaidx[i]; }
        *n_rows_out = (int64_t)sf.entries.size();
        *n_samples_out = vcf->n_samples();
        if (records_read_out) *records_read_out = nrec;
        if (index_seeks_out) *index_seeks_out = vcf->seeks();
        return NPH_OK;
    } catch (const InputError &e) {
        g_err = e.what();
        return NPH_EINPUT;
    }
}

int nph_read_gt(const char *genotype_path, uint8_t *out, int64_t row_bytes, int64_t max_records, int64_t *n_records,
                int64_t *n_samples, int32_t *width, int32_t *ploidy) {
    try {
        std::unique_ptr<VariantSource> vcf = open_variant_source(genotype_path);
        if (!vcf) return NPH_EOPEN_VCF;
        if (const char *e = getenv("NIMPRESS_READ_DS")) if (*e && *e != '0') vcf->set_dosage_mode(true);   // tests: FORMAT/DS instead of GT
        VariantRecord rec;
        int64_t k = 0;
        *n_samples = vcf->n_samples();
        while (vcf->next(rec)) {
            if (k >= max_records) return NPH_ECAPACITY;
            if (rec.has_gt) {
                vcf->load_gt(rec);
                const int64_t bytes = vcf->n_samples() * rec.ploidy * rec.gt_width;
                if (bytes > row_bytes) return NPH_ECAPACITY;
                memcpy(out + k * row_bytes, rec.gt, (size_t)bytes);
                *width = rec.gt_width; *ploidy = rec.ploidy;
            }
            k++;
        }
        *n_records = k;
        return NPH_OK;
    } catch (const InputError &e) {
        g_err = e.what();
        return NPH_EINPUT;
    }
}

int nph_fast_inflate(const uint8_t *in, int64_t in_len, uint8_t *out, int64_t out_len) {
    static thread_local FastInflateTables t;
    return fast_inflate(in, (size_t)in_len, out, (size_t)out_len, t) ? 1 : 0;
}

double nph_dbinom(int64_t x, int64_t n, double p) { return dbinom(x, n, p); }
double nph_pbinom(int64_t x, int64_t n, double p) { return pbinom(x, n, p); }
double nph_betai(double a, double b, double x) { return betai(a, b, x); }
double nph_binom_test(int64_t x, int64_t n, double p) { return binom_test(x, n, p); }

int nph_format_float(double v, char *buf, int32_t buflen) {
    std::string s = format_float_nim(v);
    if ((int32_t)s.size() + 1 > buflen) return -1;
    memcpy(buf, s.c_str(), s.size() + 1);
    return (int)s.size();
}

// ------------------------------------------------------------------------------------------
// command line -- main() of the reference (src/nimpress.nim:652-753), docopt usage text :653-706
// ------------------------------------------------------------------------------------------

static const char *USAGE =
    "Compute polygenic scores from a VCF/BCF.\n\n"
    "Usage:\n"
    "  nimpress [options] <scoredef> <genotypes.vcf>\n"
    "  nimpress [options] <scoredef1>,<scoredef2>,... <genotypes.vcf>   (one pass, a \"#score\" line per file; not in the reference)\n"
    "  nimpress (-h | --help)\n"
    "  nimpress --version\n\n"
    "Options:\n"
    "  -h --help          Show this screen.\n"
    "  --version          Show version.\n"
    "  --cov=<path>       Path to a BED file supplying genome regions that have been\n"
    "                     genotyped in the genotypes.vcf file.\n"
    "  --imp-locus=<m>    Imputation to apply for whole loci which are either not\n"
    "                     in the sequenced BED regions, or fail (too many samples\n"
    "                     with missing genotype, as set by --maxmis, or a FILTER field\n"
    "                     other than \".\" / \"PASS\" if --ignorefilt is not set). Valid\n"
    "                     values are ps, homref, fail, ignore [default: ps].\n"
    "  --imp-missing=<m>  Imputation to apply for loci which are in the sequenced BED\n"
    "                     regions (and thus should have been genotyped), but are\n"
    "                     completely missing from the VCF. Valid values are homref,\n"
    "                     ignore [default: homref].\n"
    "  --imp-sample=<m>   Imputation to apply for an individual sample with missing\n"
    "                     genotype. Valid values are ps, homref, fail, int_fail,\n"
    "                     int_ps [default: int_ps].\n"
    "  --maxmis=<f>       Maximum fraction of samples with missing genotypes allowed\n"
    "                     at a locus [default: 0.05].\n"
    "  --mincs=<n>        Minimum number of genotyped samples at a locus for internal\n"
    "                     imputation [default: 100].\n"
    "  --afmisp=<f>       p-value threshold for warning about allele frequency\n"
    "                     mismatch [default: 0.001].\n"
    "  --ignorefilt       Ignore the VCF FILTER field.\n"
    "  --device=<n>       CUDA device to score on [default: 0] (not a reference option).\n"
    "  --devices=<list>   Several CUDA devices, e.g. 0-7 or 0,2,3: the score file is split into\n"
    "                     contiguous ranges, one per device, and the partial sums are combined\n"
    "                     in range order (not a reference option).\n"
    "  --dosage           Score FORMAT/DS (expected ALT-allele dosage, one float per sample) instead of\n"
    "                     FORMAT/GT (not a reference option: the reference lists it under Future).\n"
    "  --exact-order      Add every locus to the running sums in score-file order, bit for bit\n"
    "                     like the reference (default: sum tiles of four loci first; same\n"
    "                     products, scores equal to ~1e-15 relative) (not a reference option).\n";

static int parse_enum(const std::string &v, std::initializer_list<const char *> names) {   // Nim parseEnum: style-insensitive beyond the first char is not reproduced
    int i = 0;
    for (const char *nm : names) { if (v == nm) return i; i++; }
    throw InputError("invalid enum value: " + v);
}

// "0-3", "0,2,5", "1" -> bit mask of CUDA devices
static uint32_t parse_device_list(const std::string &v) {
    uint32_t mask = 0;
    for (size_t a = 0; a <= v.size();) {
        size_t b = v.find(',', a);
        if (b == std::string::npos) b = v.size();
        const std::string item = v.substr(a, b - a);
        const size_t dash = item.find('-');
        const int64_t lo = parse_int_nim(dash == std::string::npos ? item : item.substr(0, dash), "--devices");
        const int64_t hi = dash == std::string::npos ? lo : parse_int_nim(item.substr(dash + 1), "--devices");
        if (lo < 0 || hi > 31 || hi < lo) throw InputError("--devices: device indices are 0..31, ranges lo-hi");
        for (int64_t d = lo; d <= hi; d++) mask |= 1u << d;
        a = b + 1;
    }
    if (!mask) throw InputError("--devices: empty list");
    return mask;
}

// WARN lines, then one `name<TAB>$score` line per sample (:752-753), formatted into one buffer
static void print_result(const nph_result *r) {
    std::string out = r->r.warnings;
    const std::vector<double> &s = r->r.scores;
    size_t names = 0;
    for (const std::string &nm : r->r.samples) names += nm.size();
    out.reserve(out.size() + names + s.size() * 28);
    char buf[48];
    for (size_t i = 0; i < s.size(); i++) {
        out += r->r.samples[i];
        out += '\t';
        out.append(buf, (size_t)format_float_nim_to(s[i], buf));
        out += '\n';
    }
    std::cout.write(out.data(), (std::streamsize)out.size());
}

int nph_main(int argc, char **argv) {
    nph_params p = { NPC_LOCUS_PS, NPC_MISSING_HOMREF, NPC_SAMPLE_INT_PS, 0, 0, 0, 0, 0, 100, 0.05, 0.001, 0, 0 };
    std::string cov, pos[2];
    int npos = 0;
    try {
        for (int i = 1; i < argc; i++) {
            std::string a = argv[i];
            auto val = [&](const char *opt) -> std::string {      // --opt=value or --opt value
                const size_t l = strlen(opt);
                if (a.size() > l && a[l] == '=') return a.substr(l + 1);
                if (i + 1 >= argc) throw InputError(std::string(opt) + " requires an argument");
                return argv[++i];
            };
            auto is = [&](const char *opt) { const size_t l = strlen(opt); return a.compare(0, l, opt) == 0 && (a.size() == l || a[l] == '='); };
            if (a == "-h" || a == "--help") { std::cout << USAGE; return 0; }
            else if (a == "--version") { std::cout << "nimpress 1.0.0\n"; return 0; }        // :708
            else if (is("--cov")) { cov = val("--cov"); p.use_cov = 1; }
            else if (is("--imp-locus")) p.imp_locus = parse_enum(val("--imp-locus"), { "ps", "homref", "fail", "ignore" });
            else if (is("--imp-missing")) p.imp_missing = parse_enum(val("--imp-missing"), { "homref", "ignore" });
            else if (is("--imp-sample")) p.imp_sample = parse_enum(val("--imp-sample"), { "ps", "homref", "fail", "int_ps", "int_fail" });
            else if (is("--maxmis")) p.maxmis = parse_float_nim(val("--maxmis"), "--maxmis");
            else if (is("--mincs")) p.mincs = parse_int_nim(val("--mincs"), "--mincs");
            else if (is("--afmisp")) p.afmisp = parse_float_nim(val("--afmisp"), "--afmisp");
            else if (is("--devices")) p.device_mask = (int32_t)parse_device_list(val("--devices"));
            else if (is("--device")) p.device = (int)parse_int_nim(val("--device"), "--device");
            else if (a == "--ignorefilt") p.ignorefilt = 1;
            else if (a == "--exact-order") p.exact_order = 1;
            else if (a == "--dosage") p.use_ds = 1;
            else if (a.size() > 1 && a[0] == '-') throw InputError("unknown option " + a);
            else if (npos < 2) pos[npos++] = a;
            else throw InputError("too many arguments");
        }
        if (npos != 2) throw InputError("expected <scoredef> <genotypes.vcf>");
    } catch (const InputError &e) {
        std::cerr << e.what() << "\n" << "Usage:\n  nimpress [options] <scoredef> <genotypes.vcf>\n";
        return 1;
    }
    // several score files, one pass (not a reference feature) -- unless the whole argument names a file, as it would for the reference
    if (pos[0].find(',') != std::string::npos && access(pos[0].c_str(), F_OK) != 0) {
        std::vector<std::string> paths;
        for (size_t a = 0; a <= pos[0].size();) { size_t b = pos[0].find(',', a); if (b == std::string::npos) b = pos[0].size(); if (b > a) paths.push_back(pos[0].substr(a, b - a)); a = b + 1; }
        std::vector<const char *> cp;
        for (auto &s : paths) cp.push_back(s.c_str());
        std::vector<nph_result *> rs(paths.size(), nullptr);
        int rc = nph_compute_polygenic_scores_multi(cp.data(), (int32_t)cp.size(), pos[1].c_str(), p.use_cov ? cov.c_str() : nullptr, &p, rs.data());
        if (rc == NPH_EOPEN_VCF) { std::cout << "FATAL Could not open input VCF file " << pos[1] << "\n"; return 255; }
        if (rc == NPH_EOPEN_SCORE) { std::cout << "FATAL " << nph_last_error() << "\n"; return 255; }
        if (rc) { std::cerr << "nimpress: " << nph_last_error() << "\n"; return 1; }
        for (size_t k = 0; k < rs.size(); k++) {           // per file: a "#score" line, then the reference's output for it
            std::cout << "#score\t" << paths[k] << "\n";
            print_result(rs[k]);
            nph_result_free(rs[k]);
        }
        return 0;
    }
    nph_result *r = nullptr;
    int rc = nph_compute_polygenic_scores(pos[0].c_str(), pos[1].c_str(), p.use_cov ? cov.c_str() : nullptr, &p, &r);
    if (rc == NPH_EOPEN_VCF) { std::cout << "FATAL Could not open input VCF file " << pos[1] << "\n"; return 255; }          // :729-730
    if (rc == NPH_EOPEN_SCORE) { std::cout << "FATAL Could not open polygenic score file " << pos[0] << "\n"; return 255; } // :733-734
    if (rc) { std::cerr << "nimpress: " << nph_last_error() << "\n"; return 1; }
    print_result(r);
    nph_result_free(r);
    return 0;
}
