#include "driver.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <unordered_map>

#include "stats.hpp"

namespace nph {

namespace {

struct Ctx {                                       // RAII over npc_ctx
    npc_ctx *h = nullptr;
    ~Ctx() { if (h) npc_destroy(h); }
    void ck(int rc, const char *what) const {
        if (rc != NPC_OK) throw std::runtime_error(std::string(what) + ": " + npc_last_error(h) + " (rc " + std::to_string(rc) + ")");
    }
};

// Re-encode one record's GT payload into the context layout (width w_dst, ploidy p_dst): wider
// integers with the sentinels translated, missing trailing values padded with vector_end --
// exactly the values htslib would hand the reference after widening.
bool convert_gt(const VariantRecord &rec, int64_t n, int w_dst, int p_dst, uint8_t *dst) {
    if (rec.ploidy > p_dst || rec.gt_width > w_dst) return false;
    auto load = [&](int64_t j) -> int64_t {
        if (rec.gt_width == 1) { int8_t v = ((const int8_t *)rec.gt)[j]; return v == INT8_MIN ? INT64_MIN : v == INT8_MIN + 1 ? INT64_MIN + 1 : v; }
        if (rec.gt_width == 2) { int16_t v; memcpy(&v, rec.gt + 2 * j, 2); return v == INT16_MIN ? INT64_MIN : v == INT16_MIN + 1 ? INT64_MIN + 1 : v; }
        int32_t v; memcpy(&v, rec.gt + 4 * j, 4); return v == INT32_MIN ? INT64_MIN : v == INT32_MIN + 1 ? INT64_MIN + 1 : v;
    };
    auto store = [&](int64_t j, int64_t v) {
        if (w_dst == 1) ((int8_t *)dst)[j] = v == INT64_MIN ? INT8_MIN : v == INT64_MIN + 1 ? INT8_MIN + 1 : (int8_t)v;
        else if (w_dst == 2) { int16_t t = v == INT64_MIN ? INT16_MIN : v == INT64_MIN + 1 ? INT16_MIN + 1 : (int16_t)v; memcpy(dst + 2 * j, &t, 2); }
        else { int32_t t = v == INT64_MIN ? INT32_MIN : v == INT64_MIN + 1 ? INT32_MIN + 1 : (int32_t)v; memcpy(dst + 4 * j, &t, 4); }
    };
    for (int64_t i = 0; i < n; i++)
        for (int k = 0; k < p_dst; k++) store(i * p_dst + k, k < rec.ploidy ? load(i * rec.ploidy + k) : INT64_MIN + 1);
    return true;
}

}  // namespace

Matcher::Matcher(const ScoreFile &score, const GenomeIntervals &cov, const ScoreParams &p)
    : E_(score.entries), ignorefilt_(p.ignorefilt) {
    const int64_t nE = (int64_t)E_.size();
    kind.assign(nE, PENDING); eaidx.assign(nE, -1); filter_text.assign(nE, std::string()); contig_in_bed.assign(nE, 1);
    for (int64_t i = 0; i < nE; i++) {
        if (p.use_cov) {                                            // :526, isVariantCovered :313-345
            contig_in_bed[i] = cov.has_contig(E_[i].contig);
            if (!cov.covers(E_[i])) { kind[i] = NPC_KIND_NOTCOV; continue; }   // the reference never looks these up
        }
        ContigIndex &ci = index_[E_[i].contig];
        ci.by_pos.emplace_back(E_[i].pos, i);
        ci.max_reflen = std::max<int64_t>(ci.max_reflen, (int64_t)E_[i].refseq.size());
    }
    for (auto &kv : index_) { std::sort(kv.second.by_pos.begin(), kv.second.by_pos.end()); n_lookup_ += (int64_t)kv.second.by_pos.size(); }
}

const std::vector<int64_t> &Matcher::match(const VariantRecord &rec) {
    hits_.clear();
    auto it = index_.find(*rec.contig);
    if (it == index_.end()) return hits_;
    const ContigIndex &ci = it->second;
    // entries overlapping [rec.pos, rec.end()]: entry.pos <= rec.end and entry.stop >= rec.pos
    auto lo = std::lower_bound(ci.by_pos.begin(), ci.by_pos.end(), std::make_pair(rec.pos - ci.max_reflen + 1, (int64_t)-1));
    for (auto q = lo; q != ci.by_pos.end() && q->first <= rec.end(); ++q) {
        const int64_t i = q->second;
        if (kind[i] != PENDING) continue;                           // an earlier record already matched: first one wins
        const ScoreEntry &e = E_[i];
        if (e.stop() < rec.pos || rec.ref != e.refseq) continue;
        int ea = -1;
        if (e.easeq == e.refseq) ea = 0;
        else for (size_t a = 0; a < rec.alts.size(); a++) if (rec.alts[a] == e.easeq) { ea = (int)a + 1; break; }
        if (ea < 0) continue;
        eaidx[i] = ea;
        const bool filt = !ignorefilt_ && rec.filter != "." && rec.filter != "PASS";          // :553
        kind[i] = filt ? NPC_KIND_FILTER : NPC_KIND_GT;
        if (filt) filter_text[i] = rec.filter;
        hits_.push_back(i);
    }
    return hits_;
}

void Matcher::finish() {
    for (auto &k : kind) if (k == PENDING) k = NPC_KIND_ABSENT;
}

namespace {

// NIMPRESS_TIMING=1: phase wall times on stderr
struct PhaseTimer {
    bool on = getenv("NIMPRESS_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void mark(const char *what) {
        auto now = std::chrono::steady_clock::now();
        if (on) fprintf(stderr, "[nimpress timing] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};

struct LayoutOverflow { int width, ploidy; };      // a matched record does not fit the context's GT layout

// One pass over the genotype file with a fixed GT layout.  Throws LayoutOverflow when a matched
// record needs a wider layout (the caller restarts the pass with it).
void run_pass(const ScoreFile &score, VariantSource &vcf, const GenomeIntervals &cov, const ScoreParams &p,
              int gt_width, int ploidy, ScoreResult &out) {
    const std::vector<ScoreEntry> &E = score.entries;
    const int64_t nE = (int64_t)E.size(), n = vcf.n_samples();
    out = ScoreResult();
    out.samples = vcf.samples();

    PhaseTimer timer;
    Matcher M(score, cov, p);
    timer.mark("coverage + entry index");
    std::vector<int32_t> &kind = M.kind, &eaidx = M.eaidx;
    std::vector<int64_t> slab_row(nE, -1);
    std::vector<uint8_t> done(nE, 0);
    const int64_t n_lookup = M.n_lookup();

    // staging slots of ~32 MB: big enough for efficient H2D copies, small enough that pinning three
    // of them does not dominate a short run (pinning costs ~1 ms per MB)
    const int64_t row_bytes = std::max<int64_t>(16, n * ploidy * gt_width);
    const int64_t block_rows = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(4096, (32ll << 20) / row_bytes + 1),
                                                                     std::max<int64_t>(n_lookup, 1)));
    Ctx ctx;
    npc_policy pol = { p.imp_locus, p.imp_missing, p.imp_sample, 0, p.mincs, p.maxmis };
    ctx.ck(npc_create(&ctx.h, p.device, n, ploidy, gt_width, block_rows, 3), "npc_create");
    ctx.ck(npc_set_policy(ctx.h, &pol), "npc_set_policy");
    if (p.exact_order) ctx.ck(npc_set_exact_order(ctx.h, 1), "npc_set_exact_order");
    ctx.ck(npc_reset(ctx.h), "npc_reset");
    int64_t slab_cap = 0, slab_want = std::max<int64_t>(n_lookup, 1);
    if (const char *e = getenv("NIMPRESS_SLAB_ROWS")) if (*e) slab_want = std::max<int64_t>(1, std::min<int64_t>(slab_want, atoll(e)));   // tests: force rounds
    ctx.ck(npc_resident_reserve(ctx.h, slab_want, &slab_cap), "npc_resident_reserve");
    timer.mark("GPU context + buffers");

    std::vector<int64_t> submitted;                 // entry index of every row sent to the GPU, in order
    submitted.reserve(nE);
    std::vector<npc_row> rows;
    int32_t slot = -1; uint8_t *stage = nullptr; int64_t stride = 0, staged = 0, slab_base = 0;
    auto flush_stage = [&]() {
        if (slot < 0) return;
        ctx.ck(npc_stage_upload(ctx.h, slot, staged, slab_base), "npc_stage_upload");
        slab_base += staged; staged = 0; slot = -1;
    };
    // A round scores rows over the resident slab in score-file order.  The normal case is ONE
    // final round holding every row: exactly the reference's loop order (:634-641).  Only when the
    // matched genotype rows exceed device memory are there earlier rounds; those take the rows
    // whose genotypes sit in the slab, and the summation order then differs from the reference's
    // by a reordering of terms (scores agree to ~1 ulp of the partial sums, not bit for bit).
    auto score_round = [&](bool final_round) {
        flush_stage();
        rows.clear();
        for (int64_t i = 0; i < nE; i++) {
            if (done[i]) continue;
            if (!final_round && !(kind[i] == NPC_KIND_GT && slab_row[i] >= 0)) continue;
            npc_row r;
            r.gt_row = kind[i] == NPC_KIND_GT ? (int32_t)slab_row[i] : -1;
            r.eaidx = eaidx[i]; r.beta = E[i].beta; r.eaf = E[i].eaf;
            r.ref_is_ea = E[i].ref_is_ea() ? 1 : 0; r.kind = kind[i];
            rows.push_back(r);
            submitted.push_back(i);
            done[i] = 1;
        }
        if (!rows.empty()) ctx.ck(npc_score_resident(ctx.h, rows.data(), (int64_t)rows.size()), "npc_score_resident");
        out.rounds++;
    };

    // ---- one streaming pass over the genotype file (findVariant :353-364, eaidx :375-379) ---
    VariantRecord rec;
    while (vcf.next(rec)) {
        out.records_read++;
        const std::vector<int64_t> &hits = M.match(rec);
        if (hits.empty()) continue;
        out.records_matched++;
        bool need_gt = false;
        for (int64_t i : hits) need_gt |= kind[i] == NPC_KIND_GT;
        if (!need_gt) continue;                                     // FILTER-failed: never decoded (:553-558)
        if (!rec.has_gt) throw InputError("record " + *rec.contig + ":" + std::to_string(rec.pos) + " has no GT field");
        if (rec.ploidy > ploidy || rec.gt_width > gt_width)
            throw LayoutOverflow{ std::max(rec.gt_width, gt_width), std::max(rec.ploidy, ploidy) };
        if (slab_base + staged >= slab_cap) {                       // slab full: score its rows, start over
            score_round(false);
            slab_base = 0;
        }
        if (slot < 0) {
            void *ptr;
            ctx.ck(npc_stage_acquire(ctx.h, &slot, &ptr, &stride), "npc_stage_acquire");
            stage = (uint8_t *)ptr; staged = 0;
        }
        uint8_t *dst = stage + staged * stride;
        if (rec.gt_width == gt_width && rec.ploidy == ploidy) memcpy(dst, rec.gt, (size_t)n * ploidy * gt_width);
        else convert_gt(rec, n, gt_width, ploidy, dst);
        for (int64_t i : hits) if (kind[i] == NPC_KIND_GT) slab_row[i] = slab_base + staged;
        staged++;
        if (staged == block_rows) flush_stage();
    }
    M.finish();
    timer.mark("stream + match + upload");
    score_round(true);

    // ---- results ------------------------------------------------------------------------------
    out.scores.assign(n, 0.0);
    std::vector<npc_locus> log(submitted.size());
    int64_t nlog = 0;
    ctx.ck(npc_finish(ctx.h, score.offset, out.scores.data(), &out.nloci, log.data(), (int64_t)log.size(), &nlog), "npc_finish");
    if (nlog != (int64_t)submitted.size()) throw std::runtime_error("locus log length mismatch");
    out.loci.assign(nE, npc_locus());
    for (size_t k = 0; k < submitted.size(); k++) out.loci[submitted[k]] = log[k];
    timer.mark("score + finish (GPU)");

    // ---- WARN lines, in the reference's order (:326, :527-530, :538-541, :554-557, :567-570, :575-579)
    std::string &w = out.warnings;
    for (int64_t i = 0; i < nE; i++) {
        const ScoreEntry &e = E[i];
        const npc_locus &L = out.loci[i];
        const std::string id = e.contig + ":" + std::to_string(e.pos) + ":" + e.refseq + ":" + e.easeq;
        const std::string span = e.contig + ":" + std::to_string(e.pos) + "-" + std::to_string(e.stop());
        if (L.klass == NPC_KIND_NOTCOV) {
            if (!M.contig_in_bed[i]) w += "WARN Contig " + e.contig + " not present within the coverage BED file.\n";
            w += "WARN Locus " + span + " is not covered by the sequence coverage BED.  Imputing all dosages at this locus.\n";
        } else if (L.klass == NPC_KIND_ABSENT) {
            if (!std::isnan(e.eaf) && binom_test(0, n * 2, e.eaf) < p.afmisp)
                w += "WARN Variant " + id + " cohort EAF is 0 in " + std::to_string(n) + " samples.  This is highly unlikely given polygenic score EAF of " +
                     format_float_nim(e.eaf) + "\n";
        } else if (L.klass == NPC_KIND_FILTER) {
            w += "WARN Variant " + id + " has a FILTER flag set (value \"" + M.filter_text[i] + "\").  Imputing all dosages at this locus.\n";
        } else if (L.klass == NPC_CLASS_MAXMIS) {
            const double missingrate = (double)L.nmiss / (double)n;
            w += "WARN Locus " + span + " has " + format_float_nim(missingrate * 100) +
                 "% of samples missing a genotype. This exceeds the missingness threshold; imputing all dosages at this locus.\n";
        } else if (!std::isnan(e.eaf) && binom_test(L.neff, (n - L.nmiss) * 2, e.eaf) < p.afmisp) {
            w += "WARN Variant " + id + " cohort EAF is " + format_float_nim((double)L.neff / (double)((n - L.nmiss) * 2)) + " in " +
                 std::to_string(n) + " samples.  This is highly unlikely given polygenic score EAF of " + format_float_nim(e.eaf) + "\n";
        }
    }
    timer.mark("WARN lines (binomial tests)");
}

}  // namespace

bool compute_polygenic_scores(const ScoreFile &score, const std::string &genotype_path, const GenomeIntervals &cov,
                              const ScoreParams &p, ScoreResult &out) {
    int width = 1, ploidy = 2;                      // BCF's usual GT layout: int8, diploid
    for (int attempt = 0; attempt < 4; attempt++) {
        std::unique_ptr<VariantSource> vcf = open_variant_source(genotype_path);
        if (!vcf) return false;
        try {
            run_pass(score, *vcf, cov, p, width, ploidy, out);
            return true;
        } catch (const LayoutOverflow &o) {         // rare: wider integers or higher ploidy -- rerun with that layout
            width = o.width; ploidy = o.ploidy;
        }
    }
    throw std::runtime_error("GT layout kept growing between passes");
}

}  // namespace nph
