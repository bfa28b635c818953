#include "driver.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <future>
#include <map>
#include <mutex>
#include <thread>
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

void Matcher::add_spans(std::unordered_map<std::string, std::vector<std::pair<int64_t, int64_t>>> &spans) const {
    for (const auto &kv : index_) {
        auto &v = spans[kv.first];
        for (const auto &q : kv.second.by_pos) v.emplace_back(E_[q.second].pos, E_[q.second].stop());
    }
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

// CUDA driver + primary-context creation is most of a short run's start-up (~1 s per process on the boxes seen):
// it is started once per device on a thread of its own as soon as a scoring call begins -- while the inputs are
// opened, indexed and, when there is a .tbi / .csi, consulted -- and waited for only where a context is needed.
std::mutex g_warm_mutex;
std::map<int, std::shared_future<void>> g_warm;
void start_warm(int device) {
    std::lock_guard<std::mutex> g(g_warm_mutex);
    if (!g_warm.count(device)) g_warm[device] = std::async(std::launch::async, [device] { npc_warmup(device); }).share();
}
void wait_warm(int device) {
    std::shared_future<void> f;
    {
        std::lock_guard<std::mutex> g(g_warm_mutex);
        auto it = g_warm.find(device);
        if (it == g_warm.end()) return;
        f = it->second;
    }
    f.wait();
}

struct LayoutOverflow { int width, ploidy; };      // a matched record does not fit the context's GT layout
struct SlabOverflow {};                            // several score files, and their matched rows exceed device memory
struct NoIndex {};                                 // the index-driven pass is not possible / not worthwhile: stream the file

// WARN lines of one score file, in the reference's order (:326, :527-530, :538-541, :554-557, :567-570, :575-579)
std::string make_warnings(const ScoreFile &score, const Matcher &M, const std::vector<npc_locus> &loci, int64_t n, const ScoreParams &p) {
    std::string w;
    const std::vector<ScoreEntry> &E = score.entries;
    for (int64_t i = 0; i < (int64_t)E.size(); i++) {
        const ScoreEntry &e = E[i];
        const npc_locus &L = loci[i];
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
        } else if (p.use_ds) {
            // dosage rows: the tally is a real number, the reference's exact binomial test is not defined on it
        } else if (!std::isnan(e.eaf) && binom_test(L.neff, (n - L.nmiss) * 2, e.eaf) < p.afmisp) {
            w += "WARN Variant " + id + " cohort EAF is " + format_float_nim((double)L.neff / (double)((n - L.nmiss) * 2)) + " in " +
                 std::to_string(n) + " samples.  This is highly unlikely given polygenic score EAF of " + format_float_nim(e.eaf) + "\n";
        }
    }
    return w;
}

// One pass over the genotype file with a fixed GT layout, for one score file or several at once
// (BASELINE config 4: every score file sees the same record stream, a record any of them matches
// is uploaded once, and the resident slab is then scored for each file).  Throws LayoutOverflow
// when a matched record needs a wider layout (the caller restarts the pass with it) and
// SlabOverflow when several files are given and their rows do not fit device memory.
void run_pass(const std::vector<const ScoreFile *> &scores, VariantSource &vcf, const std::string &genotype_path, bool seekable,
              const GenomeIntervals &cov, const ScoreParams &p, int gt_width, int ploidy, std::vector<ScoreResult> &outs) {
    const int S = (int)scores.size();
    const int64_t n = vcf.n_samples();
    struct PerScore {
        std::unique_ptr<Matcher> M;
        std::vector<int64_t> slab_row;              // per entry: row of its device's resident slab
        std::vector<uint8_t> done;
        std::vector<npc_row> rows;
    };
    std::vector<PerScore> ps(S);
    outs.assign(S, ScoreResult());

    PhaseTimer timer;
    int64_t n_lookup = 0;
    {   // entries with the same (contig, pos, ref, ea) settle on the same record: count keys once
        std::unordered_map<std::string, int> keys;
        for (int k = 0; k < S; k++) {
            outs[k].samples = vcf.samples();
            ps[k].M.reset(new Matcher(*scores[k], cov, p));
            const int64_t nE = (int64_t)scores[k]->entries.size();
            ps[k].slab_row.assign(nE, -1); ps[k].done.assign(nE, 0);
            if (S == 1) { n_lookup = ps[k].M->n_lookup(); break; }
            for (int64_t i = 0; i < nE; i++) if (ps[k].M->kind[i] == Matcher::PENDING) {
                const ScoreEntry &e = scores[k]->entries[i];
                keys.emplace(e.contig + "\t" + std::to_string(e.pos) + "\t" + e.refseq + "\t" + e.easeq, 0);
            }
        }
        if (S > 1) n_lookup = (int64_t)keys.size();
    }
    if (seekable) {                                   // an index next to the file and few loci: read only where loci are
        std::unordered_map<std::string, std::vector<std::pair<int64_t, int64_t>>> spans;
        for (int k = 0; k < S; k++) ps[k].M->add_spans(spans);
        if (!vcf.use_regions(spans, genotype_path)) throw NoIndex();
    }
    timer.mark("coverage + entry index");

    // Devices: one context per GPU.  With several, the score rows are partitioned into contiguous ranges in
    // score-file order (a contig / region partition for a sorted score file): context d scores range d over all
    // samples, a matched record is uploaded to the device(s) whose ranges name it, and the partial sums are
    // combined in range order by npc_reduce.  Several score files at once: every file is cut into D ranges of its own,
    // device d scores range d of every file in one npc_score_resident_multi call (raw partial sums), and the host adds
    // the partials of each file in device order and normalises once.
    std::vector<int> dev_ids = p.devices.empty() ? std::vector<int>{ p.device } : p.devices;
    if (const char *e = getenv("NIMPRESS_SPLIT")) if (*e && dev_ids.size() == 1 && atoi(e) > 1)      // tests: several contexts on one GPU
        dev_ids.assign((size_t)std::min(atoi(e), 16), dev_ids[0]);
    const int D = (int)dev_ids.size();
    // entry i of score file k belongs to device i * D / (entries of k): contiguous ranges of EVERY file in its own order
    auto dev_of = [&](int k, int64_t i) -> int {
        return D == 1 ? 0 : (int)std::min<int64_t>(D - 1, i * D / std::max<int64_t>((int64_t)scores[k]->entries.size(), 1));
    };
    std::vector<int64_t> dev_lookup(D, 0);            // rows of each device's ranges that need a record (an upper bound when files share records)
    if (D == 1) dev_lookup[0] = n_lookup;
    else for (int k = 0; k < S; k++)
        for (int64_t i = 0; i < (int64_t)scores[k]->entries.size(); i++) if (ps[k].M->kind[i] == Matcher::PENDING) dev_lookup[dev_of(k, i)]++;

    // staging slots of ~8 MB: big enough for efficient H2D copies, small enough that pinning three of them per
    // device does not dominate a short run (pinning costs ~1 ms per MB); scoring launches are sized separately
    const int64_t row_bytes = std::max<int64_t>(16, n * ploidy * gt_width);
    struct Dev {
        Ctx ctx;
        int32_t slot = -1; uint8_t *stage = nullptr; int64_t stride = 0, staged = 0, slab_base = 0, slab_cap = 0, block_rows = 1;
        std::vector<npc_row> rows; std::vector<int64_t> submitted;   // S == 1: this device's rows, entry index of each
        std::string err;
    };
    std::vector<Dev> devs(D);
    npc_policy pol = { p.imp_locus, p.imp_missing, p.imp_sample, 0, p.mincs, p.maxmis };
    auto open_dev = [&](int d) {
        Dev &v = devs[d];
        try {
            const int64_t want = std::max<int64_t>(dev_lookup[d], 1);
            v.block_rows = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(4096, (8ll << 20) / row_bytes + 1), want));
            const int64_t launch_rows = std::max<int64_t>(v.block_rows, 32768);
            v.ctx.ck(npc_create2(&v.ctx.h, dev_ids[d], n, ploidy, gt_width, launch_rows, 3, v.block_rows), "npc_create");
            v.ctx.ck(npc_set_policy(v.ctx.h, &pol), "npc_set_policy");
            if (p.use_ds) v.ctx.ck(npc_set_dosage_rows(v.ctx.h, 1), "npc_set_dosage_rows");
            if (p.exact_order) v.ctx.ck(npc_set_exact_order(v.ctx.h, 1), "npc_set_exact_order");
            v.ctx.ck(npc_reset(v.ctx.h), "npc_reset");
            int64_t slab_want = want;
            if (const char *e = getenv("NIMPRESS_SLAB_ROWS")) if (*e) slab_want = std::max<int64_t>(1, std::min<int64_t>(slab_want, atoll(e)));   // tests: force rounds
            v.ctx.ck(npc_resident_reserve(v.ctx.h, slab_want, &v.slab_cap), "npc_resident_reserve");
        } catch (const std::exception &e) { v.err = e.what(); }
    };
    // A device is opened when its first row arrives (or at the end, for its constant rows): contexts come up one after
    // the other inside the driver (~1 s each), and a sorted file reaches the later ranges later -- the first device's
    // rows stream while the others' contexts are still being created.
    std::vector<uint8_t> opened(D, 0);
    double waited_ms = 0.0;
    auto ensure_open = [&](int d) {
        if (opened[d]) return;
        const auto t0 = std::chrono::steady_clock::now();
        wait_warm(dev_ids[d]);
        open_dev(d);
        waited_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (!devs[d].err.empty()) throw std::runtime_error(devs[d].err);
        opened[d] = 1;
    };
    ensure_open(0);
    timer.mark("first GPU context + buffers");

    auto flush_stage = [&](Dev &v) {
        if (v.slot < 0) return;
        v.ctx.ck(npc_stage_upload(v.ctx.h, v.slot, v.staged, v.slab_base), "npc_stage_upload");
        v.slab_base += v.staged; v.staged = 0; v.slot = -1;
    };
    // the rows of score k that a round takes on device d, in score-file order
    auto collect_rows = [&](int k, int d, bool final_round, std::vector<npc_row> &rows, std::vector<int64_t> &submitted) {
        PerScore &q = ps[k];
        const std::vector<ScoreEntry> &E = scores[k]->entries;
        rows.clear();
        for (int64_t i = 0; i < (int64_t)E.size(); i++) {
            if (q.done[i] || dev_of(k, i) != d) continue;
            const int32_t kind = q.M->kind[i];
            if (!final_round && !(kind == NPC_KIND_GT && q.slab_row[i] >= 0)) continue;
            npc_row r;
            r.gt_row = kind == NPC_KIND_GT ? (int32_t)q.slab_row[i] : -1;
            r.eaidx = q.M->eaidx[i]; r.beta = E[i].beta; r.eaf = E[i].eaf;
            r.ref_is_ea = E[i].ref_is_ea() ? 1 : 0; r.kind = kind;
            rows.push_back(r);
            submitted.push_back(i);
            q.done[i] = 1;
        }
    };
    // A round scores rows over a device's resident slab in score-file order.  The normal case is ONE
    // final round per device holding every row of its range: exactly the reference's loop order
    // (:634-641).  Only when the matched genotype rows exceed device memory are there earlier rounds;
    // those take the rows whose genotypes sit in the slab, and the summation order then differs from
    // the reference's by a reordering of terms (scores agree to ~1 ulp of the partial sums, not bit for bit).
    std::vector<int64_t> rounds(D, 0);
    auto score_round = [&](int d, bool final_round) {     // one score file only
        Dev &v = devs[d];
        flush_stage(v);
        collect_rows(0, d, final_round, v.rows, v.submitted);
        if (!v.rows.empty()) v.ctx.ck(npc_score_resident(v.ctx.h, v.rows.data(), (int64_t)v.rows.size()), "npc_score_resident");
        rounds[d]++;
    };

    // ---- one streaming pass over the genotype file (findVariant :353-364, eaidx :375-379) ---
    VariantRecord rec;
    int64_t records_read = 0, records_matched = 0;
    std::vector<int64_t> rec_row(D);                   // slab row this record got on each device, -1 = not uploaded there
    while (vcf.next(rec)) {
        records_read++;
        bool any = false, need_gt = false;
        uint32_t need_dev = 0;
        for (int k = 0; k < S; k++) {
            const std::vector<int64_t> &hits = ps[k].M->match(rec);
            for (int64_t i : hits) {
                any = true;
                if (ps[k].M->kind[i] == NPC_KIND_GT) { need_gt = true; need_dev |= 1u << dev_of(k, i); }
            }
        }
        if (!any) continue;
        records_matched++;
        if (!need_gt) continue;                                     // FILTER-failed: never decoded (:553-558)
        const uint8_t *first = nullptr;
        for (int d = 0; d < D; d++) {
            rec_row[d] = -1;
            if (!((need_dev >> d) & 1u)) continue;
            ensure_open(d);
            Dev &v = devs[d];
            if (v.slab_base + v.staged >= v.slab_cap) {             // slab full: score its rows, start over
                if (S > 1) throw SlabOverflow();
                score_round(d, false);
                v.slab_base = 0;
            }
            if (v.slot < 0) {
                void *ptr;
                v.ctx.ck(npc_stage_acquire(v.ctx.h, &v.slot, &ptr, &v.stride), "npc_stage_acquire");
                v.stage = (uint8_t *)ptr; v.staged = 0;
            }
            uint8_t *dst = v.stage + v.staged * v.stride;
            if (first) memcpy(dst, first, (size_t)n * ploidy * gt_width);   // a record named by two ranges: same bytes
            else {
                // BCF with the context's layout: the GT payload goes from the inflated blocks straight into the pinned row
                const bool direct = vcf.load_gt_into(rec, dst, gt_width, ploidy);
                if (!rec.has_gt || !rec.gt)
                    throw InputError("record " + *rec.contig + ":" + std::to_string(rec.pos) + (p.use_ds ? " has no DS field" : " has no GT field"));
                if (p.use_ds && (rec.ploidy != 1 || rec.gt_width != 4))
                    throw InputError("record " + *rec.contig + ":" + std::to_string(rec.pos) + ": FORMAT/DS with several values per sample is not supported");
                if (!direct) {
                    if (rec.ploidy > ploidy || rec.gt_width > gt_width)
                        throw LayoutOverflow{ std::max(rec.gt_width, gt_width), std::max(rec.ploidy, ploidy) };
                    if (rec.gt_width == gt_width && rec.ploidy == ploidy) memcpy(dst, rec.gt, (size_t)n * ploidy * gt_width);
                    else convert_gt(rec, n, gt_width, ploidy, dst);
                }
                first = dst;
            }
            rec_row[d] = v.slab_base + v.staged;
            v.staged++;
        }
        for (int k = 0; k < S; k++) for (int64_t i : ps[k].M->match_last())
            if (ps[k].M->kind[i] == NPC_KIND_GT) {
                if (p.use_ds && ps[k].M->eaidx[i] > 1)
                    throw InputError("record " + *rec.contig + ":" + std::to_string(rec.pos) + ": FORMAT/DS rows take the REF or the first ALT as effect allele");
                ps[k].slab_row[i] = rec_row[dev_of(k, i)];
            }
        for (int d = 0; d < D; d++) if (devs[d].slot >= 0 && devs[d].staged == devs[d].block_rows) flush_stage(devs[d]);
    }
    for (int k = 0; k < S; k++) {
        ps[k].M->finish();
        outs[k].records_read = records_read; outs[k].records_matched = records_matched; outs[k].index_seeks = vcf.seeks();
    }
    timer.mark("stream + match + upload");
    if (timer.on) fprintf(stderr, "[nimpress timing] records read %lld, matched %lld, index seeks %lld%s\n", (long long)records_read,
                          (long long)records_matched, (long long)vcf.seeks(), vcf.seeks() ? " (.tbi / .csi used)" : "");

    // ---- score the slab(s), collect results ----------------------------------------------------
    std::vector<std::vector<npc_locus>> logs(S);
    std::vector<std::vector<int64_t>> order(S);          // entry index of every log record
    for (int d = 0; d < D; d++) ensure_open(d);
    if (timer.on && D > 1) fprintf(stderr, "[nimpress timing] waited for GPU contexts + buffers, all devices %8.1f ms\n", waited_ms);
    if (S == 1) {
        for (int d = 0; d < D; d++) score_round(d, true);
        outs[0].scores.assign(n, 0.0);
        outs[0].rounds = *std::max_element(rounds.begin(), rounds.end());
        outs[0].devices = D;
        if (D == 1) {
            Dev &v = devs[0];
            logs[0].resize(v.submitted.size());
            int64_t nlog = 0;
            v.ctx.ck(npc_finish(v.ctx.h, scores[0]->offset, outs[0].scores.data(), &outs[0].nloci, logs[0].data(), (int64_t)logs[0].size(), &nlog), "npc_finish");
            if (nlog != (int64_t)v.submitted.size()) throw std::runtime_error("locus log length mismatch");
            order[0] = v.submitted;
        } else {
            std::vector<npc_ctx *> hs(D);
            for (int d = 0; d < D; d++) hs[d] = devs[d].ctx.h;
            const double off = scores[0]->offset;
            devs[0].ctx.ck(npc_reduce(hs.data(), D, &off, outs[0].scores.data(), &outs[0].nloci), "npc_reduce");
            std::vector<double> scratch((size_t)std::max<int64_t>(n, 1));
            for (int d = 0; d < D; d++) {
                Dev &v = devs[d];
                std::vector<npc_locus> lg(v.submitted.size());
                int64_t nlog = 0, nl = 0;
                v.ctx.ck(npc_partial(v.ctx.h, scratch.data(), &nl, lg.data(), (int64_t)lg.size(), &nlog), "npc_partial");
                if (nlog != (int64_t)v.submitted.size()) throw std::runtime_error("locus log length mismatch");
                logs[0].insert(logs[0].end(), lg.begin(), lg.end());
                order[0].insert(order[0].end(), v.submitted.begin(), v.submitted.end());
            }
        }
    } else {
        std::vector<int64_t> nloci_tot(S, 0);
        std::vector<std::vector<double>> part(S);
        for (int k = 0; k < S; k++) { outs[k].scores.assign(n, 0.0); outs[k].rounds = 1; outs[k].devices = D; if (D > 1) part[k].assign((size_t)n, 0.0); }
        for (int d = 0; d < D; d++) {
            Dev &v = devs[d];
            flush_stage(v);
            std::vector<std::vector<npc_row>> drows(S);
            std::vector<std::vector<npc_locus>> dlog(S);
            std::vector<const npc_row *> rows(S);
            std::vector<int64_t> n_rows(S), nloci(S);
            std::vector<double> offsets(S);
            std::vector<double *> sc(S);
            std::vector<npc_locus *> lg(S);
            for (int k = 0; k < S; k++) {
                collect_rows(k, d, true, drows[k], order[k]);
                rows[k] = drows[k].data(); n_rows[k] = (int64_t)drows[k].size(); offsets[k] = scores[k]->offset;
                sc[k] = D == 1 ? outs[k].scores.data() : part[k].data();
                dlog[k].resize(drows[k].size()); lg[k] = dlog[k].data();
            }
            // one device: normalised scores straight away; several: raw partial sums (offsets = NULL), combined below
            v.ctx.ck(npc_score_resident_multi(v.ctx.h, S, rows.data(), n_rows.data(), D == 1 ? offsets.data() : nullptr, sc.data(), nloci.data(), lg.data()),
                     "npc_score_resident_multi");
            for (int k = 0; k < S; k++) {
                nloci_tot[k] += nloci[k];
                logs[k].insert(logs[k].end(), dlog[k].begin(), dlog[k].end());
                if (D > 1) {
                    double *tot = outs[k].scores.data();
                    const double *pk = part[k].data();
                    if (d == 0) memcpy(tot, pk, (size_t)n * sizeof(double));
                    else for (int64_t i = 0; i < n; i++) tot[i] += pk[i];            // device order = score-file order of the ranges
                }
            }
        }
        for (int k = 0; k < S; k++) {
            outs[k].nloci = nloci_tot[k];
            if (D > 1) npc_normalise(outs[k].scores.data(), n, nloci_tot[k], scores[k]->offset);
        }
    }
    timer.mark("score + finish (GPU)");
    for (int k = 0; k < S; k++) {
        outs[k].loci.assign(scores[k]->entries.size(), npc_locus());
        for (size_t j = 0; j < order[k].size(); j++) outs[k].loci[order[k][j]] = logs[k][j];
        outs[k].warnings = make_warnings(*scores[k], *ps[k].M, outs[k].loci, n, p);
    }
    timer.mark("WARN lines (binomial tests)");
}

}  // namespace

bool compute_polygenic_scores_multi(const std::vector<const ScoreFile *> &scores, const std::string &genotype_path,
                                    const GenomeIntervals &cov, const ScoreParams &p, std::vector<ScoreResult> &outs) {
    for (int d : (p.devices.empty() ? std::vector<int>{ p.device } : p.devices)) start_warm(d);
    int width = 1, ploidy = 2;                      // BCF's usual GT layout: int8, diploid
    if (p.use_ds) { width = 4; ploidy = 1; }        // FORMAT/DS: one float per sample
    bool seekable = true;                           // first try the index-driven pass (needs <file>.tbi / .csi)
    for (int attempt = 0; attempt < 6; attempt++) {
        std::unique_ptr<VariantSource> vcf = open_variant_source(genotype_path, seekable);
        if (!vcf) return false;
        vcf->set_dosage_mode(p.use_ds);
        try {
            run_pass(scores, *vcf, genotype_path, seekable, cov, p, width, ploidy, outs);
            return true;
        } catch (const NoIndex &) {
            seekable = false;
        } catch (const LayoutOverflow &o) {         // rare: wider integers or higher ploidy -- rerun with that layout
            width = o.width; ploidy = o.ploidy;
        } catch (const SlabOverflow &) {            // too many rows for one resident slab: one file at a time
            outs.assign(scores.size(), ScoreResult());
            for (size_t k = 0; k < scores.size(); k++)
                if (!compute_polygenic_scores(*scores[k], genotype_path, cov, p, outs[k])) return false;
            return true;
        }
    }
    throw std::runtime_error("GT layout kept growing between passes");
}

bool compute_polygenic_scores(const ScoreFile &score, const std::string &genotype_path, const GenomeIntervals &cov,
                              const ScoreParams &p, ScoreResult &out) {
    std::vector<ScoreResult> outs;
    if (!compute_polygenic_scores_multi({ &score }, genotype_path, cov, p, outs)) return false;
    out = std::move(outs[0]);
    return true;
}

}  // namespace nph
