#include "variant_source.hpp"
#include "region_index.hpp"
#include "fast_inflate.hpp"

#include <algorithm>
#include <atomic>
#include <climits>
#include <condition_variable>
#include <cstring>
#include <map>
#include <mutex>
#include <thread>
#include <unordered_map>

namespace nph {

// ------------------------------------------------------------------------------------------
// InflateStream
// ------------------------------------------------------------------------------------------

// ------------------------------------------------------------------------------------------
// BgzfPool: a reader thread cuts the file into BGZF blocks (18-byte header, BSIZE in the `BC`
// extra field), worker threads inflate them (raw deflate, independent streams), the consumer takes
// them in file order from a ring of slots.
// ------------------------------------------------------------------------------------------

class BgzfPool {
public:
    BgzfPool(FILE *fp, int n_workers) : fp_(fp), ring_((size_t)n_workers * 4) {
        reader_ = std::thread([this] { read_loop(); });
        for (int i = 0; i < n_workers; i++) workers_.emplace_back([this] { work_loop(); });
    }
    ~BgzfPool() {
        { std::lock_guard<std::mutex> g(m_); stop_ = true; }
        cv_job_.notify_all(); cv_free_.notify_all(); cv_done_.notify_all();
        reader_.join();
        for (auto &t : workers_) t.join();
    }
    // next inflated block, in order; false at end of file.  The returned buffer stays valid until the next call.
    bool next(const uint8_t *&data, size_t &len) {
        std::unique_lock<std::mutex> g(m_);
        if (held_) { ring_[(rd_ - 1) % ring_.size()].state = FREE; held_ = false; cv_free_.notify_all(); }
        for (;;) {
            Slot &s = ring_[rd_ % ring_.size()];
            if (rd_ < wr_ && s.state == DONE) {
                if (!s.error.empty()) throw InputError(s.error);
                data = s.out.data(); len = s.out_len; rd_++; held_ = true;
                return true;
            }
            if (eof_ && rd_ == wr_) { if (!error_.empty()) throw InputError(error_); return false; }
            cv_done_.wait(g);
        }
    }
private:
    enum { FREE = 0, QUEUED = 1, BUSY = 2, DONE = 3 };
    struct Slot { std::vector<uint8_t> in, out; size_t in_len = 0, out_len = 0; int state = FREE; std::string error; };
    void read_loop() {
        for (;;) {
            size_t idx;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_free_.wait(g, [this] { return stop_ || ring_[wr_ % ring_.size()].state == FREE; });
                if (stop_) return;
                idx = wr_ % ring_.size();
            }
            Slot &s = ring_[idx];
            uint8_t hdr[18];
            size_t got = fread(hdr, 1, 18, fp_);
            std::string err;
            size_t bsize = 0;
            if (got == 0) { finish(""); return; }
            if (got != 18 || hdr[0] != 0x1f || hdr[1] != 0x8b || !(hdr[3] & 4) || hdr[12] != 'B' || hdr[13] != 'C')
                err = "BGZF: malformed block header";
            else {
                bsize = (size_t)(hdr[16] | (hdr[17] << 8)) + 1;
                const size_t xlen = (size_t)(hdr[10] | (hdr[11] << 8));
                if (xlen < 6 || bsize < 12 + xlen + 8) err = "BGZF: bad block size";      // xlen < 6: no room for the BC subfield we just read
                else {
                    const size_t rest = bsize - 18;                 // remaining extra fields + deflate data + CRC32 + ISIZE
                    if (s.in.size() < rest + 16) s.in.resize(std::max<size_t>(rest + 16, (1 << 16) + 64));   // + the decoder's read-ahead
                    if (fread(s.in.data(), 1, rest, fp_) != rest) err = "BGZF: truncated block";
                    else {
                        const size_t skip = xlen - 6;               // extra subfields beyond BC
                        s.in_len = rest; s.out_len = skip;          // out_len doubles as the payload offset for the worker
                    }
                }
            }
            if (!err.empty()) { finish(err); return; }
            {
                std::lock_guard<std::mutex> g(m_);
                s.state = QUEUED; s.error.clear(); wr_++;
            }
            cv_job_.notify_one();
        }
    }
    void finish(const std::string &err) {
        std::lock_guard<std::mutex> g(m_);
        eof_ = true; error_ = err;
        cv_done_.notify_all(); cv_job_.notify_all();
    }
    void work_loop() {
        z_stream zs;
        memset(&zs, 0, sizeof zs);
        inflateInit2(&zs, -15);
        std::unique_ptr<FastInflateTables> tabs(new FastInflateTables);
        const char *zo = getenv("NIMPRESS_ZLIB_ONLY");
        const bool zlib_only = zo && *zo && *zo != '0';
        for (;;) {
            Slot *s = nullptr;
            {
                std::unique_lock<std::mutex> g(m_);
                for (;;) {
                    if (stop_) { inflateEnd(&zs); return; }
                    for (size_t k = job_; k < wr_; k++)
                        if (ring_[k % ring_.size()].state == QUEUED) { s = &ring_[k % ring_.size()]; s->state = BUSY; if (k == job_) job_++; break; }
                    if (s) break;
                    while (job_ < wr_ && ring_[job_ % ring_.size()].state != QUEUED) job_++;
                    if (job_ < wr_) continue;
                    if (eof_) { inflateEnd(&zs); return; }
                    cv_job_.wait(g);
                }
            }
            const size_t off = s->out_len, tail = 8;
            uint32_t isize, want_crc;
            memcpy(&want_crc, s->in.data() + s->in_len - 8, 4);
            memcpy(&isize, s->in.data() + s->in_len - 4, 4);
            if (s->out.size() < (size_t)isize + 16) s->out.resize(std::max<size_t>((size_t)isize + 16, (1 << 16) + 16));
            std::string err;
            // this engine's own decoder first (fast_inflate.hpp); zlib when it declines or the CRC disagrees
            bool ok = !zlib_only && isize && s->in_len >= off + tail &&
                      fast_inflate(s->in.data() + off, s->in_len - off - tail, s->out.data(), isize, *tabs) &&
                      (uint32_t)crc32(0L, s->out.data(), isize) == want_crc;
            if (!ok && isize) {
                inflateReset(&zs);
                zs.next_in = s->in.data() + off; zs.avail_in = (uInt)(s->in_len - off - tail);
                zs.next_out = s->out.data(); zs.avail_out = (uInt)s->out.size();
                const int rc = inflate(&zs, Z_FINISH);
                if (rc != Z_STREAM_END || zs.total_out != isize) err = "BGZF: corrupt deflate data";
                else if ((uint32_t)crc32(0L, s->out.data(), isize) != want_crc) err = "BGZF: CRC32 mismatch";
            }
            {
                std::lock_guard<std::mutex> g(m_);
                s->out_len = isize; s->error = err; s->state = DONE;
            }
            cv_done_.notify_all();
        }
    }
    FILE *fp_;
    std::vector<Slot> ring_;
    std::mutex m_;
    std::condition_variable cv_job_, cv_free_, cv_done_;
    size_t wr_ = 0, rd_ = 0, job_ = 0;
    bool eof_ = false, stop_ = false, held_ = false;
    std::string error_;
    std::thread reader_;
    std::vector<std::thread> workers_;
};

InflateStream::InflateStream() = default;

InflateStream::~InflateStream() {
    pool_.reset();
    if (zinit_) inflateEnd(&zs_);
    if (fp_) fclose(fp_);
}

bool InflateStream::open(const std::string &path, bool seekable) {
    fp_ = fopen(path.c_str(), "rb");
    if (!fp_) return false;
    fseek(fp_, 0, SEEK_END);
    file_size_ = (int64_t)ftell(fp_);
    fseek(fp_, 0, SEEK_SET);
    unsigned char magic[18] = { 0 };
    size_t got = fread(magic, 1, 18, fp_);
    fseek(fp_, 0, SEEK_SET);
    compressed_ = got >= 2 && magic[0] == 0x1f && magic[1] == 0x8b;
    const bool bgzf = got == 18 && compressed_ && (magic[3] & 4) && magic[12] == 'B' && magic[13] == 'C';
    bgzf_ = bgzf;
    // all cores but one (the reader's own thread walks the records and copies the GT payloads), at most 32
    const unsigned hc = std::max(1u, std::thread::hardware_concurrency());
    int want = (int)std::min<unsigned>(hc > 2 ? hc - 1 : hc, 32u);
    if (const char *e = getenv("NIMPRESS_THREADS")) if (*e) want = std::max(1, atoi(e));
    block_mode_ = bgzf && (seekable || want == 1);     // one thread: block by block with this engine's decoder
    if (bgzf && want > 1 && !block_mode_) {
        threads_ = want;
        pool_ = std::make_unique<BgzfPool>(fp_, want);
        return true;
    }
    in_.resize(block_mode_ ? 128 << 10 : 1 << 20);    // seekable: a jump costs one read of two blocks, not a megabyte
    out_.resize(4 << 20);
    if (compressed_) {
        memset(&zs_, 0, sizeof zs_);
        if (inflateInit2(&zs_, 15 + 16) != Z_OK) return false;
        zinit_ = true;
    }
    return true;
}

bool InflateStream::fill() {
    out_pos_ = out_len_ = 0;
    if (eof_) return false;
    if (pool_) {                                      // blocks arrive inflated, in order; empty blocks (EOF marker) are skipped
        const uint8_t *d; size_t l;
        do {
            if (!pool_->next(d, l)) { eof_ = true; return false; }
        } while (l == 0);
        pool_data_ = d; out_len_ = l;
        return true;
    }
    if (!compressed_) {
        out_len_ = fread(out_.data(), 1, out_.size(), fp_);
        if (out_len_ == 0) eof_ = true;
        return out_len_ > 0;
    }
    if (block_mode_) {                                // exactly one BGZF block per fill: the position stays a virtual offset
        for (;;) {
            block_coff_ = in_file_off_;
            uint8_t hdr[18];
            const size_t got = fread(hdr, 1, 18, fp_);
            if (got == 0) { eof_ = true; return false; }
            if (got != 18 || hdr[0] != 0x1f || hdr[1] != 0x8b || !(hdr[3] & 4) || hdr[12] != 'B' || hdr[13] != 'C')
                throw InputError("BGZF: malformed block header");
            const size_t bsize = (size_t)(hdr[16] | (hdr[17] << 8)) + 1, xlen = (size_t)(hdr[10] | (hdr[11] << 8));
            if (bsize < 12 + xlen + 8 || xlen < 6) throw InputError("BGZF: bad block size");
            const size_t rest = bsize - 18, off = xlen - 6;         // remaining extra fields, deflate data, CRC32, ISIZE
            if (in_.size() < rest + 16) in_.resize(rest + 16);
            if (fread(in_.data(), 1, rest, fp_) != rest) throw InputError("BGZF: truncated block");
            in_file_off_ += bsize;
            uint32_t want_crc, isize;
            memcpy(&want_crc, in_.data() + rest - 8, 4);
            memcpy(&isize, in_.data() + rest - 4, 4);
            if (isize == 0) continue;                               // empty blocks (the EOF marker) are skipped
            if (out_.size() < (size_t)isize + 16) out_.resize((size_t)isize + 16);
            if (!tabs_) tabs_.reset(new FastInflateTables);
            const char *zo = getenv("NIMPRESS_ZLIB_ONLY");
            bool ok = !(zo && *zo && *zo != '0') && rest >= off + 8 &&
                      fast_inflate(in_.data() + off, rest - off - 8, out_.data(), isize, *tabs_) &&
                      (uint32_t)crc32(0L, out_.data(), isize) == want_crc;
            if (!ok) {                                              // zlib, raw deflate
                if (inflateReset2(&zs_, -15) != Z_OK) throw InputError("zlib: inflateReset failed");
                zs_.next_in = in_.data() + off; zs_.avail_in = (uInt)(rest - off - 8);
                zs_.next_out = out_.data(); zs_.avail_out = (uInt)out_.size();
                const int rc = inflate(&zs_, Z_FINISH);
                if (rc != Z_STREAM_END || zs_.total_out != isize) throw InputError("BGZF: corrupt deflate data");
                if ((uint32_t)crc32(0L, out_.data(), isize) != want_crc) throw InputError("BGZF: CRC32 mismatch");
            }
            out_len_ = isize;
            return true;
        }
    }
    zs_.next_out = out_.data();
    zs_.avail_out = (uInt)out_.size();
    while (zs_.avail_out == out_.size()) {            // until something was produced
        if (zs_.avail_in == 0) {
            zs_.next_in = in_.data();
            zs_.avail_in = (uInt)fread(in_.data(), 1, in_.size(), fp_);
            if (zs_.avail_in == 0) { eof_ = true; break; }
        }
        int rc = inflate(&zs_, Z_NO_FLUSH);
        if (rc == Z_STREAM_END) {                     // next BGZF block / gzip member
            if (inflateReset(&zs_) != Z_OK) throw InputError("zlib: inflateReset failed");
        } else if (rc != Z_OK && rc != Z_BUF_ERROR) {
            throw InputError(std::string("zlib: corrupt compressed stream: ") + (zs_.msg ? zs_.msg : "?"));
        }
    }
    out_len_ = out_.size() - zs_.avail_out;
    return out_len_ > 0;
}

void InflateStream::seek_virtual(uint64_t voff) {
    if (!block_mode_) throw InputError("seek on a stream that was not opened seekable");
    const uint64_t coff = voff >> 16;
    if (fseek(fp_, (long)coff, SEEK_SET) != 0) throw InputError("index points outside the file");
    in_file_off_ = coff;
    eof_ = false;
    out_pos_ = out_len_ = 0;
    if (!fill()) return;                              // at or past the end: the next read reports end of stream
    out_pos_ = std::min<size_t>((size_t)(voff & 0xFFFF), out_len_);
}

size_t InflateStream::read(void *dst, size_t n) {
    size_t done = 0;
    while (done < n) {
        if (out_pos_ == out_len_ && !fill()) break;
        size_t k = std::min(n - done, out_len_ - out_pos_);
        memcpy((uint8_t *)dst + done, cur() + out_pos_, k);
        out_pos_ += k;
        done += k;
    }
    return done;
}

bool InflateStream::read_exact(void *dst, size_t n) { return read(dst, n) == n; }

bool InflateStream::skip(size_t n) {                  // advance without copying
    while (n) {
        if (out_pos_ == out_len_ && !fill()) return false;
        const size_t k = std::min(n, out_len_ - out_pos_);
        out_pos_ += k; n -= k;
    }
    return true;
}

bool InflateStream::peek(void *dst, size_t n) {
    if (out_pos_ == out_len_ && !fill()) return false;
    if (out_len_ - out_pos_ < n) return false;        // only used at the very start of the stream
    memcpy(dst, cur() + out_pos_, n);
    return true;
}

bool InflateStream::getline(std::string &line) {
    line.clear();
    bool any = false;
    for (;;) {
        if (out_pos_ == out_len_ && !fill()) break;
        any = true;
        const uint8_t *b = cur() + out_pos_;
        const uint8_t *nl = (const uint8_t *)memchr(b, '\n', out_len_ - out_pos_);
        if (nl) {
            line.append((const char *)b, nl - b);
            out_pos_ += (nl - b) + 1;
            break;
        }
        line.append((const char *)b, out_len_ - out_pos_);
        out_pos_ = out_len_;
    }
    if (!any) return false;
    if (!line.empty() && line.back() == '\r') line.pop_back();
    return true;
}

// ------------------------------------------------------------------------------------------
// shared: encode GT values into the narrowest BCF integer width, pad with vector_end
// ------------------------------------------------------------------------------------------

static void pack_gt(const std::vector<int32_t> &vals, const std::vector<int> &counts, int64_t n, int ploidy,
                    std::vector<uint8_t> &out, int &width) {
    int32_t mx = 0;
    for (int32_t v : vals) mx = std::max(mx, v);
    width = mx <= INT8_MAX ? 1 : mx <= INT16_MAX ? 2 : 4;
    out.assign((size_t)n * ploidy * width, 0);
    size_t src = 0;
    for (int64_t i = 0; i < n; i++) {
        for (int k = 0; k < ploidy; k++) {
            const bool have = k < counts[i];
            const int32_t v = have ? vals[src + k] : 0;
            const size_t j = (size_t)i * ploidy + k;
            if (width == 1) ((int8_t *)out.data())[j] = have ? (int8_t)v : (int8_t)(INT8_MIN + 1);
            else if (width == 2) ((int16_t *)out.data())[j] = have ? (int16_t)v : (int16_t)(INT16_MIN + 1);
            else ((int32_t *)out.data())[j] = have ? v : INT32_MIN + 1;
        }
        src += counts[i];
    }
}

// INFO/END=<n> overrides the REF length for overlap (htslib vcf_parse / tabix readrec)
static int64_t rlen_from_info(const char *info, size_t len, int64_t pos, int64_t dflt) {
    size_t i = 0;
    while (i < len) {
        size_t e = i;
        while (e < len && info[e] != ';') e++;
        if (e - i > 4 && !memcmp(info + i, "END=", 4)) {
            long long end = atoll(std::string(info + i + 4, e - i - 4).c_str());
            return end >= pos ? end - pos + 1 : dflt;
        }
        i = e + 1;
    }
    return dflt;
}

// ------------------------------------------------------------------------------------------
// VCF text
// ------------------------------------------------------------------------------------------

class VcfTextSource : public VariantSource {
public:
    explicit VcfTextSource(std::unique_ptr<InflateStream> s) : in_(std::move(s)) {}
    bool read_header() {
        while (in_->getline(line_)) {
            if (line_.size() >= 2 && line_[0] == '#' && line_[1] == '#') continue;
            if (!line_.empty() && line_[0] == '#') {
                std::vector<std::string> f = split_char(line_, '\t');
                for (size_t i = 9; i < f.size(); i++) samples_.push_back(f[i]);
                return true;
            }
            return false;                              // data before the #CHROM line
        }
        return false;
    }
    InflateStream *stream() override { return in_.get(); }
    bool index_names_contigs() const override { return index_ && !index_->names().empty(); }
    int contig_rank(const VariantRecord &rec) const override { return contig_rank(*rec.contig); }
    int contig_rank(const std::string &name) const override {
        if (!index_) return -1;
        if (rank_.empty()) for (size_t i = 0; i < index_->names().size(); i++) rank_.emplace(index_->names()[i], (int)i);
        auto it = rank_.find(name);
        return it == rank_.end() ? -1 : it->second;
    }
    bool next_raw(VariantRecord &rec) override {
        for (;;) {
            if (!in_->getline(line_)) return false;
            if (!line_.empty()) break;
        }
        // split the 9 fixed columns in place
        const char *p = line_.data(), *end = p + line_.size();
        const char *col[10];
        size_t len[10];
        int nc = 0;
        while (nc < 9 && p <= end) {
            const char *t = (const char *)memchr(p, '\t', end - p);
            col[nc] = p; len[nc] = (t ? t : end) - p; nc++;
            if (!t) { p = end + 1; break; }
            p = t + 1;
        }
        if (nc < 8) throw InputError("VCF: record with fewer than 8 columns");
        contig_.assign(col[0], len[0]);
        rec.contig = &contig_; rec.contig_id = -1;
        rec.pos = parse_int_nim(std::string(col[1], len[1]), "VCF POS");
        rec.ref.assign(col[3], len[3]);
        rec.alts.clear();
        if (!(len[4] == 1 && col[4][0] == '.')) rec.alts = split_char(std::string(col[4], len[4]), ',');
        rec.filter.assign(col[6], len[6]);
        rec.rlen = rlen_from_info(col[7], len[7], rec.pos, (int64_t)rec.ref.size());
        rec.has_gt = false; rec.gt = nullptr; rec.ploidy = 0; rec.gt_width = 1;
        if (nc < 9 || samples_.empty()) return true;
        // FORMAT: index of GT
        int gt_idx = -1, k = 0;
        for (const char *q = col[8], *qe = col[8] + len[8]; q <= qe; k++) {
            const char *t = (const char *)memchr(q, ':', qe - q);
            size_t l = (t ? t : qe) - q;
            if (l == 2 && (dosage_ ? (q[0] == 'D' && q[1] == 'S') : (q[0] == 'G' && q[1] == 'T'))) gt_idx = k;
            if (!t) break;
            q = t + 1;
        }
        if (gt_idx < 0) return true;
        // the genotype columns are parsed only if the caller asks (load_gt): most records of a
        // genome-wide file match no score row
        rec.has_gt = true;
        gt_p_ = p; gt_idx_ = gt_idx; gt_loaded_ = false;
        return true;
    }
    void load_gt(VariantRecord &rec) override {
        if (!rec.has_gt || gt_loaded_) return;
        gt_loaded_ = true;
        const char *p = gt_p_, *end = line_.data() + line_.size();
        const int64_t n = n_samples();
        if (dosage_) {                                   // DS sub-field of every sample -> BCF floats ("." = missing)
            gt_.resize((size_t)n * 4);
            int maxv = 1;
            for (int64_t i = 0; i < n; i++) {
                const char *s = p <= end ? p : end, *se = end;
                if (p <= end) {
                    const char *t = (const char *)memchr(p, '\t', end - p);
                    se = t ? t : end;
                    p = t ? t + 1 : end + 1;
                }
                for (int f = 0; f < gt_idx_ && s < se; f++) {
                    const char *t = (const char *)memchr(s, ':', se - s);
                    s = t ? t + 1 : se;
                }
                const char *fe = (const char *)memchr(s, ':', se - s);
                if (!fe) fe = se;
                if (memchr(s, ',', fe - s)) maxv = 2;             // Number=A with several ALTs: the caller refuses
                uint32_t bits = 0x7F800001u;
                if (s < fe && !(fe - s == 1 && *s == '.')) {
                    const float v = (float)parse_float_nim(std::string(s, fe - s), "FORMAT/DS");
                    memcpy(&bits, &v, 4);
                }
                memcpy(gt_.data() + 4 * i, &bits, 4);
            }
            rec.ploidy = maxv; rec.gt_width = 4; rec.gt = gt_.data();
            return;
        }
        // fast path: every sample is "a/b" or "a|b" with one-character alleles and GT the first
        // sub-field -- written straight as int8 (allele+1)<<1|phased (htslib vcf_parse_format)
        if (gt_idx_ == 0 && p <= end) {
            gt_.resize((size_t)n * 2);
            int8_t *out = (int8_t *)gt_.data();
            const char *s = p;
            int64_t i = 0;
            for (; i < n; i++) {
                if (end - s < 3) break;
                const char c0 = s[0], sep = s[1], c1 = s[2];
                const bool ok0 = c0 == '.' || (c0 >= '0' && c0 <= '9'), ok1 = c1 == '.' || (c1 >= '0' && c1 <= '9');
                if (!ok0 || !ok1 || (sep != '/' && sep != '|')) break;
                const char *t = s + 3;
                if (t < end && *t != '\t') {
                    if (*t != ':') break;                          // longer allele number or another ploidy
                    t = (const char *)memchr(t, '\t', end - t);
                    if (!t) t = end;
                }
                out[2 * i] = c0 == '.' ? 0 : (int8_t)((c0 - '0' + 1) << 1);
                out[2 * i + 1] = (int8_t)((c1 == '.' ? 0 : ((c1 - '0' + 1) << 1)) | (sep == '|'));
                s = t + 1;
                if (t >= end && i + 1 < n) { i++; break; }
            }
            if (i == n) { rec.ploidy = 2; rec.gt_width = 1; rec.gt = gt_.data(); return; }
        }
        // general path: per-sample GT strings -> (allele+1)<<1|phased, any ploidy, any allele number
        vals_.clear(); counts_.assign(n, 0);
        int maxp = 1;
        const int gt_idx = gt_idx_;
        for (int64_t i = 0; i < n; i++) {
            const char *s = p <= end ? p : end, *se = end;
            if (p <= end) {
                const char *t = (const char *)memchr(p, '\t', end - p);
                se = t ? t : end;
                p = t ? t + 1 : end + 1;
            }
            for (int f = 0; f < gt_idx && s < se; f++) {          // skip to sub-field gt_idx
                const char *t = (const char *)memchr(s, ':', se - s);
                s = t ? t + 1 : se;
            }
            int l = 0, phased = 0;
            for (;;) {
                if (s < se && *s == '.') { s++; vals_.push_back(phased); l++; }
                else if (s < se && *s >= '0' && *s <= '9') {
                    uint32_t v = 0;
                    while (s < se && *s >= '0' && *s <= '9') v = v * 10 + (uint32_t)(*s++ - '0');
                    vals_.push_back((int32_t)(((v + 1) << 1) | (uint32_t)phased)); l++;
                } else break;
                if (s >= se) break;
                phased = *s == '|';
                if (*s != '|' && *s != '/') break;
                s++;
            }
            if (!l) { vals_.push_back(0); l = 1; }                // empty field = one missing allele
            counts_[i] = l;
            maxp = std::max(maxp, l);
        }
        pack_gt(vals_, counts_, n, maxp, gt_, rec.gt_width);
        rec.ploidy = maxp; rec.gt = gt_.data();
    }
private:
    std::unique_ptr<InflateStream> in_;
    std::string line_, contig_;
    const char *gt_p_ = nullptr;
    int gt_idx_ = -1;
    bool gt_loaded_ = true;
    mutable std::unordered_map<std::string, int> rank_;
    std::vector<int32_t> vals_;
    std::vector<int> counts_;
    std::vector<uint8_t> gt_;
};

// ------------------------------------------------------------------------------------------
// BCF2
// ------------------------------------------------------------------------------------------

class BcfSource : public VariantSource {
public:
    explicit BcfSource(std::unique_ptr<InflateStream> s) : in_(std::move(s)) {}
    bool read_header() {
        uint8_t magic[5];
        uint32_t l_text = 0;
        if (!in_->read_exact(magic, 5) || memcmp(magic, "BCF\2", 4) != 0) return false;
        if (!in_->read_exact(&l_text, 4)) return false;
        std::string text(l_text, '\0');
        if (!in_->read_exact(&text[0], l_text)) return false;
        parse_header(text);
        return true;
    }
    InflateStream *stream() override { return in_.get(); }
    void after_seek() override { indiv_left_ = 0; gt_done_ = true; }
    int contig_rank(const VariantRecord &rec) const override { return rec.contig_id; }       // CSI reference ids = header contig ids
    int contig_rank(const std::string &name) const override {
        for (size_t i = 0; i < contigs_.size(); i++) if (contigs_[i] == name) return (int)i;
        return -1;
    }
    bool next_raw(VariantRecord &rec) override {
        // the per-sample block of the previous record is consumed only on demand (load_gt / load_gt_into):
        // a record no score row matches costs no copy of its genotypes at all
        if (indiv_left_ && !in_->skip(indiv_left_)) throw InputError("BCF: truncated record");
        indiv_left_ = 0;
        uint32_t lens[2];
        size_t got = in_->read(lens, 8);
        if (got == 0) return false;
        if (got != 8) throw InputError("BCF: truncated record header");
        shared_.resize(lens[0]);
        if (!in_->read_exact(shared_.data(), lens[0])) throw InputError("BCF: truncated record");
        indiv_left_ = lens[1];
        if (lens[0] < 24) throw InputError("BCF: shared block too short");
        const uint8_t *p = shared_.data(), *e = p + shared_.size();
        int32_t chrom, pos0, rlen; uint32_t nai, nfs;
        memcpy(&chrom, p, 4); memcpy(&pos0, p + 4, 4); memcpy(&rlen, p + 8, 4);
        memcpy(&nai, p + 16, 4); memcpy(&nfs, p + 20, 4);
        p += 24;
        const uint32_t n_allele = nai >> 16, n_info = nai & 0xFFFF, n_fmt = nfs >> 24, n_sample = nfs & 0xFFFFFF;
        (void)n_info;
        if (chrom < 0 || (size_t)chrom >= contigs_.size() || contigs_[chrom].empty()) throw InputError("BCF: CHROM id not in header");
        rec.contig_id = chrom; rec.contig = &contigs_[chrom];
        rec.pos = (int64_t)pos0 + 1; rec.rlen = rlen;
        skip_typed(p, e);                                          // ID
        rec.ref.clear(); rec.alts.clear();
        for (uint32_t a = 0; a < n_allele; a++) {
            std::string s = read_typed_string(p, e);
            if (a == 0) rec.ref = std::move(s); else rec.alts.push_back(std::move(s));
        }
        // FILTER: typed int vector of dictionary ids
        {
            int type; uint32_t len;
            read_desc(p, e, type, len);
            rec.filter.clear();
            if (len == 0) rec.filter = ".";
            for (uint32_t i = 0; i < len; i++) {
                int32_t id = read_int(p, e, type);
                if (i) rec.filter += ';';
                rec.filter += dict_name(id);
            }
        }
        // FORMAT fields are looked at when the genotypes are asked for
        rec.has_gt = n_fmt > 0 && n_sample > 0; rec.gt = nullptr; rec.ploidy = 0; rec.gt_width = 1;
        n_fmt_ = n_fmt; n_sample_ = n_sample; gt_done_ = false;
        return true;
    }
    // whole per-sample block into indiv_, GT located inside it
    void load_gt(VariantRecord &rec) override {
        if (gt_done_) return;
        gt_done_ = true;
        indiv_.resize(indiv_left_);
        if (!in_->read_exact(indiv_.data(), indiv_left_)) throw InputError("BCF: truncated record");
        indiv_left_ = 0;
        rec.has_gt = false; rec.gt = nullptr; rec.ploidy = 0; rec.gt_width = 1;
        const uint8_t *q = indiv_.data(), *qe = q + indiv_.size();
        for (uint32_t f = 0; f < n_fmt_ && q < qe; f++) {
            int kt; uint32_t kl;
            read_desc(q, qe, kt, kl);
            int32_t key = read_int(q, qe, kt);
            int type; uint32_t len;
            read_desc(q, qe, type, len);
            const size_t esz = type == 1 ? 1 : type == 2 ? 2 : type == 3 ? 4 : type == 5 ? 4 : type == 7 ? 1 : 0;
            const size_t bytes = (size_t)n_sample_ * len * esz;
            if (q + bytes > qe) throw InputError("BCF: FORMAT field overruns the record");
            if (wanted(key, type)) {
                if ((int64_t)n_sample_ != n_samples()) throw InputError("BCF: record sample count differs from header");
                rec.has_gt = true; rec.gt = q; rec.gt_width = (int)esz; rec.ploidy = (int)len;
            }
            q += bytes;
        }
    }
    // The GT payload straight from the inflated blocks into dst (the pinned staging row) when its layout is the wanted
    // one: the only copy the genotypes make on the host.  The field headers in front of it are a few bytes each.
    bool load_gt_into(VariantRecord &rec, uint8_t *dst, int want_width, int want_ploidy) override {
        if (gt_done_) return false;
        uint8_t hdr[16];
        size_t left = indiv_left_;
        for (uint32_t f = 0; f < n_fmt_ && left; f++) {
            // key (typed int) and the type descriptor: read what a short header needs, byte by byte where lengths chain
            size_t h = 0;
            auto need = [&](size_t n) { if (left < n || !in_->read_exact(hdr + h, n)) throw InputError("BCF: truncated record"); h += n; left -= n; };
            need(1);
            const int kt = hdr[0] & 0xF;
            if ((hdr[0] >> 4) != 1 || (kt != 1 && kt != 2 && kt != 3)) { finish_general(rec, hdr, h, left); return false; }
            const size_t ksz = kt == 1 ? 1 : kt == 2 ? 2 : 4;
            need(ksz);
            const uint8_t *kp = hdr + 1;
            const int32_t key = read_int(kp, hdr + h, kt);
            const size_t d0 = h;
            need(1);
            int type = hdr[d0] & 0xF; uint32_t len = hdr[d0] >> 4;
            if (len == 15) {                                                // long vector: the length follows as a typed int
                need(1);
                const int lt = hdr[d0 + 1] & 0xF;
                if ((hdr[d0 + 1] >> 4) != 1 || (lt != 1 && lt != 2 && lt != 3)) { finish_general(rec, hdr, h, left); return false; }
                need(lt == 1 ? 1 : lt == 2 ? 2 : 4);
                const uint8_t *lp = hdr + d0 + 2;
                len = (uint32_t)read_int(lp, hdr + h, lt);
            }
            const size_t esz = type == 1 ? 1 : type == 2 ? 2 : type == 3 ? 4 : type == 5 ? 4 : type == 7 ? 1 : 0;
            const size_t bytes = (size_t)n_sample_ * len * esz;
            if (bytes > left) throw InputError("BCF: FORMAT field overruns the record");
            if (wanted(key, type)) {
                if ((int64_t)n_sample_ != n_samples()) throw InputError("BCF: record sample count differs from header");
                gt_done_ = true;
                rec.has_gt = true; rec.gt_width = (int)esz; rec.ploidy = (int)len;
                const bool direct = (int)esz == want_width && (int)len == want_ploidy;
                uint8_t *to = dst;
                if (!direct) { indiv_.resize(bytes); to = indiv_.data(); }
                if (!in_->read_exact(to, bytes)) throw InputError("BCF: truncated record");
                rec.gt = to;
                indiv_left_ = left - bytes;                                 // later fields: skipped when the next record is read
                return direct;
            }
            if (!in_->skip(bytes)) throw InputError("BCF: truncated record");
            left -= bytes;
        }
        indiv_left_ = left; gt_done_ = true;
        rec.has_gt = false; rec.gt = nullptr;
        return false;
    }
private:
    static void read_desc(const uint8_t *&p, const uint8_t *e, int &type, uint32_t &len) {
        if (p >= e) throw InputError("BCF: truncated typed value");
        const uint8_t d = *p++;
        type = d & 0xF; len = d >> 4;
        if (len == 15) {
            int t2; uint32_t l2;
            read_desc(p, e, t2, l2);
            len = (uint32_t)read_int(p, e, t2);
        }
    }
    static int32_t read_int(const uint8_t *&p, const uint8_t *e, int type) {
        int32_t v = 0;
        if (type == 1) { if (p + 1 > e) goto bad; v = (int8_t)*p; p += 1; }
        else if (type == 2) { if (p + 2 > e) goto bad; int16_t t; memcpy(&t, p, 2); v = t; p += 2; }
        else if (type == 3) { if (p + 4 > e) goto bad; memcpy(&v, p, 4); p += 4; }
        else throw InputError("BCF: expected an integer typed value");
        return v;
    bad:
        throw InputError("BCF: truncated integer");
    }
    static void skip_typed(const uint8_t *&p, const uint8_t *e) {
        int type; uint32_t len;
        read_desc(p, e, type, len);
        const size_t esz = type == 1 ? 1 : type == 2 ? 2 : type == 3 ? 4 : type == 5 ? 4 : type == 7 ? 1 : 0;
        if (p + (size_t)len * esz > e) throw InputError("BCF: truncated typed value");
        p += (size_t)len * esz;
    }
    static std::string read_typed_string(const uint8_t *&p, const uint8_t *e) {
        int type; uint32_t len;
        read_desc(p, e, type, len);
        if (type != 7 && len != 0) throw InputError("BCF: expected a string typed value");
        if (p + len > e) throw InputError("BCF: truncated string");
        std::string s((const char *)p, len);
        p += len;
        while (!s.empty() && s.back() == '\0') s.pop_back();
        return s;
    }
    const std::string &dict_name(int32_t id) const {
        static const std::string unknown = "?";
        return id >= 0 && (size_t)id < dict_.size() && !dict_[id].empty() ? dict_[id] : unknown;
    }
    // ID=..., optional IDX=... of one ##KEY=<...> line
    static bool meta_id(const std::string &line, size_t lt, std::string &id, int &idx) {
        idx = -1;
        size_t p = line.find("ID=", lt);
        if (p == std::string::npos) return false;
        size_t e = line.find_first_of(",>", p);
        id = line.substr(p + 3, (e == std::string::npos ? line.size() : e) - p - 3);
        size_t q = line.find("IDX=", lt);
        if (q != std::string::npos) idx = atoi(line.c_str() + q + 4);
        return true;
    }
    void parse_header(const std::string &text) {
        std::unordered_map<std::string, int> seen;
        auto add_dict = [&](const std::string &id, int idx) {
            auto it = seen.find(id);
            if (it != seen.end()) return;
            if (idx < 0) idx = (int)dict_next_++;
            else dict_next_ = std::max<size_t>(dict_next_, (size_t)idx + 1);
            if (dict_.size() <= (size_t)idx) dict_.resize(idx + 1);
            dict_[idx] = id;
            seen[id] = idx;
        };
        add_dict("PASS", 0);
        size_t b = 0;
        while (b < text.size()) {
            size_t e = text.find('\n', b);
            if (e == std::string::npos) e = text.size();
            std::string line = text.substr(b, e - b);
            while (!line.empty() && (line.back() == '\r' || line.back() == '\0')) line.pop_back();
            b = e + 1;
            if (line.rfind("##", 0) == 0) {
                size_t eq = line.find("=<");
                if (eq == std::string::npos) continue;
                const std::string key = line.substr(2, eq - 2);
                std::string id; int idx;
                if (!meta_id(line, eq, id, idx)) continue;
                if (key == "contig") {
                    if (idx < 0) idx = (int)contig_next_++;
                    else contig_next_ = std::max<size_t>(contig_next_, (size_t)idx + 1);
                    if (contigs_.size() <= (size_t)idx) contigs_.resize(idx + 1);
                    contigs_[idx] = id;
                } else if (key == "FILTER" || key == "INFO" || key == "FORMAT") {
                    add_dict(id, idx);
                }
            } else if (!line.empty() && line[0] == '#') {
                std::vector<std::string> f = split_char(line, '\t');
                for (size_t i = 9; i < f.size(); i++) samples_.push_back(f[i]);
            }
        }
        auto it = seen.find("GT");
        gt_key_ = it == seen.end() ? -1 : it->second;
        it = seen.find("DS");
        ds_key_ = it == seen.end() ? -1 : it->second;
    }
    // the FORMAT field the caller reads: GT (typed ints) or, in dosage mode, DS (floats)
    bool wanted(int32_t key, int type) const {
        return dosage_ ? (key == ds_key_ && type == 5) : (key == gt_key_ && (type == 1 || type == 2 || type == 3));
    }
    std::unique_ptr<InflateStream> in_;
    std::vector<std::string> contigs_, dict_;
    size_t dict_next_ = 0, contig_next_ = 0;
    // an unusual field header inside load_gt_into: put the bytes read so far back in front and take the general path
    void finish_general(VariantRecord &rec, const uint8_t *hdr, size_t h, size_t left) {
        indiv_.resize(h + left);
        memcpy(indiv_.data(), hdr, h);
        if (!in_->read_exact(indiv_.data() + h, left)) throw InputError("BCF: truncated record");
        // fields already skipped are gone; the general parser starts at this field
        indiv_left_ = 0; gt_done_ = true;
        rec.has_gt = false; rec.gt = nullptr; rec.ploidy = 0; rec.gt_width = 1;
        const uint8_t *q = indiv_.data(), *qe = q + indiv_.size();
        while (q < qe) {
            int kt; uint32_t kl;
            read_desc(q, qe, kt, kl);
            int32_t key = read_int(q, qe, kt);
            int type; uint32_t len;
            read_desc(q, qe, type, len);
            const size_t esz = type == 1 ? 1 : type == 2 ? 2 : type == 3 ? 4 : type == 5 ? 4 : type == 7 ? 1 : 0;
            const size_t bytes = (size_t)n_sample_ * len * esz;
            if (q + bytes > qe) throw InputError("BCF: FORMAT field overruns the record");
            if (wanted(key, type)) { rec.has_gt = true; rec.gt = q; rec.gt_width = (int)esz; rec.ploidy = (int)len; }
            q += bytes;
        }
    }
    int32_t gt_key_ = -1, ds_key_ = -1;
    size_t indiv_left_ = 0;
    uint32_t n_fmt_ = 0, n_sample_ = 0;
    bool gt_done_ = true;
    std::vector<uint8_t> shared_, indiv_;
};

// ------------------------------------------------------------------------------------------
// index-driven region mode
// ------------------------------------------------------------------------------------------

VariantSource::VariantSource() = default;
VariantSource::~VariantSource() = default;

bool VariantSource::use_regions(const std::unordered_map<std::string, std::vector<std::pair<int64_t, int64_t>>> &spans,
                                const std::string &path) {
    if (const char *e = getenv("NIMPRESS_NO_INDEX")) if (*e && *e != '0') return false;
    InflateStream *in = stream();
    if (!in || !in->is_bgzf()) return false;
    index_ = RegionIndex::load(path);
    if (!index_) return false;
    if (!index_names_contigs()) { index_.reset(); return false; }
    std::vector<Region> regs;
    for (const auto &kv : spans) {
        const int ref = contig_rank(kv.first);
        if (ref < 0 || ref >= index_->n_ref()) continue;            // contig not in the file: its loci stay unmatched, as when streaming
        for (const auto &sp : kv.second) regs.push_back(Region{ ref, sp.first - 1, sp.second });
    }
    std::sort(regs.begin(), regs.end(), [](const Region &a, const Region &b) { return a.ref != b.ref ? a.ref < b.ref : a.beg0 < b.beg0; });
    std::vector<Region> merged;                                       // spans closer than one 16 kb index window are one region
    for (const Region &r : regs) {
        if (!merged.empty() && merged.back().ref == r.ref && r.beg0 <= merged.back().end0 + (1 << 14)) merged.back().end0 = std::max(merged.back().end0, r.end0);
        else merged.push_back(r);
    }
    // Jumping reads block by block with one zlib stream (no inflate pool: several times slower per byte than
    // streaming).  Estimate the compressed bytes the regions need from the index itself -- from the first record
    // of a region to the first record at its end, plus a block -- and use the index only when that is a small
    // part of the file.
    const char *force = getenv("NIMPRESS_FORCE_INDEX");
    if (!(force && *force && *force != '0')) {
        int64_t est = 0;
        for (const Region &r : merged) {
            const uint64_t v0 = index_->query_start(r.ref, r.beg0, r.end0);
            if (v0 == RegionIndex::NONE) continue;
            const uint64_t v1 = index_->query_start(r.ref, r.end0, r.end0 + 1);
            const int64_t c0 = (int64_t)(v0 >> 16), c1 = v1 == RegionIndex::NONE ? in->file_size() : (int64_t)(v1 >> 16);
            est += std::max<int64_t>(c1 - c0, 0) + (64 << 10);
        }
        if (est * 8 > in->file_size()) { index_.reset(); return false; }
    }
    // The pass below walks the regions in header-contig order and only ever seeks forward.  htslib merely
    // requires each contig's records to be contiguous, so a file may hold its contig blocks in another order
    // than its header names them (and index fine): the regions' first records must then be visited in FILE
    // order, which the rank-ordered walk cannot do -- stream such a file instead of silently missing loci.
    {
        uint64_t last = 0;
        for (const Region &r : merged) {
            const uint64_t v0 = index_->query_start(r.ref, r.beg0, r.end0);
            if (v0 == RegionIndex::NONE) continue;
            if (v0 < last) { index_.reset(); return false; }
            last = v0;
        }
    }
    regions_ = std::move(merged);
    region_ = 0; filtering_ = true; positioned_ = false;
    return true;
}

bool VariantSource::next(VariantRecord &rec) {
    if (!filtering_) return next_raw(rec);
    InflateStream *in = stream();
    for (;;) {
        if (region_ >= regions_.size()) return false;
        if (!positioned_) {                                           // entering regions_[region_]: jump if its first record lies ahead
            const Region &R = regions_[region_];
            const uint64_t v = index_->query_start(R.ref, R.beg0, R.end0);
            if (v == RegionIndex::NONE) { region_++; continue; }
            if (v > in->tell_virtual()) { in->seek_virtual(v); after_seek(); seeks_++; }
            positioned_ = true;
        }
        if (!next_raw(rec)) return false;
        for (;;) {                                                    // place the record against the current region
            const Region &R = regions_[region_];
            const int rank = contig_rank(rec);
            const bool past = rank > R.ref || (rank == R.ref && rec.pos - 1 >= R.end0);
            if (!past) {
                if (rank == R.ref && rec.end() > R.beg0) return true; // overlaps (rec.end() is 1-based inclusive = 0-based exclusive)
                break;                                                // before the region (or a contig the index does not know): keep reading
            }
            if (++region_ >= regions_.size()) return false;
            const Region &N = regions_[region_];
            const uint64_t v = index_->query_start(N.ref, N.beg0, N.end0);
            if (v == RegionIndex::NONE) { positioned_ = false; break; }
            if (v > in->tell_virtual()) { in->seek_virtual(v); after_seek(); seeks_++; positioned_ = true; break; }   // this record lies before the next region's first
            positioned_ = true;                                       // no jump: the same record may belong to the next region
        }
        if (!positioned_) continue;
    }
}

std::unique_ptr<VariantSource> open_variant_source(const std::string &path, bool seekable) {
    auto s = std::make_unique<InflateStream>();
    if (!s->open(path, seekable)) return nullptr;
    uint8_t magic[5] = { 0 };
    try {
        if (!s->peek(magic, 5)) return nullptr;
        if (!memcmp(magic, "BCF\2", 4)) {
            auto src = std::make_unique<BcfSource>(std::move(s));
            if (!src->read_header()) return nullptr;
            return src;
        }
        if (magic[0] != '#') return nullptr;
        auto src = std::make_unique<VcfTextSource>(std::move(s));
        if (!src->read_header()) return nullptr;
        return src;
    } catch (const InputError &) {
        return nullptr;
    }
}

}  // namespace nph
