// variant_source.hpp -- streaming reader of VCF text / BCF2, plain, gzip or BGZF, written against
// zlib only.  It stands in for what the reference gets from hts-nim/htslib (third-party, not in
// the reference tree): sample names, and per record CHROM, POS, rlen, REF, ALT, FILTER and the
// FORMAT/GT payload.  For BCF the GT payload is handed out as the raw typed-value bytes of the
// record -- exactly the slab the CUDA kernels decode -- without widening or per-sample parsing;
// for VCF text the GT strings are encoded the way htslib's vcf_parse_format encodes them.
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <memory>
#include <string>
#include <vector>

#include "util.hpp"

namespace nph {

class BgzfPool;                                      // multi-threaded BGZF block inflater (variant_source.cpp)

// Byte stream over a plain / gzip / BGZF file.  BGZF (concatenated <= 64 KiB gzip members with a
// `BC` extra field giving each member's size) is inflated by a pool of worker threads, blocks
// delivered in file order; plain gzip falls back to one sequential zlib stream.  This is the
// replacement for htslib's bgzf reader (which the reference uses single-threaded, hts-nim's
// default threads = 0).  NIMPRESS_THREADS sets the pool size (default: min(cores, 16); 1 = sequential).
class InflateStream {
public:
    InflateStream();
    ~InflateStream();
    bool open(const std::string &path);
    size_t read(void *dst, size_t n);                 // up to n bytes; 0 at end of stream
    bool read_exact(void *dst, size_t n);
    bool getline(std::string &line);                  // strips \n and a preceding \r
    bool peek(void *dst, size_t n);                   // look ahead without consuming (n <= 64)
    int threads() const { return threads_; }
private:
    bool fill();
    FILE *fp_ = nullptr;
    bool compressed_ = false, zinit_ = false, eof_ = false;
    z_stream zs_{};
    std::vector<uint8_t> in_, out_;
    size_t out_pos_ = 0, out_len_ = 0;
    const uint8_t *cur() const { return pool_ ? pool_data_ : out_.data(); }
    std::unique_ptr<BgzfPool> pool_;
    const uint8_t *pool_data_ = nullptr;
    int threads_ = 1;
};

struct VariantRecord {
    int32_t contig_id = -1;
    const std::string *contig = nullptr;
    int64_t pos = 0;                                  // 1-based
    int64_t rlen = 0;                                 // reference span used for overlap (REF length or INFO/END)
    std::string ref;
    std::vector<std::string> alts;
    std::string filter;                               // ".", "PASS" or ';'-joined names
    bool has_gt = false;
    int gt_width = 1, ploidy = 0;                     // bytes per value, values per sample            } after
    const uint8_t *gt = nullptr;                      // n_samples * ploidy * gt_width bytes, until next() } load_gt()
    int64_t end() const { return pos + rlen - 1; }    // 1-based inclusive
};

class VariantSource {
public:
    virtual ~VariantSource() = default;
    virtual bool next(VariantRecord &rec) = 0;        // false at end of file; throws InputError on corrupt input
    // Fills rec.gt / ploidy / gt_width of the record next() just returned (rec.has_gt tells whether
    // it has a GT field).  Text VCF parses its genotype columns only here; BCF has them already.
    virtual void load_gt(VariantRecord &) {}
    const std::vector<std::string> &samples() const { return samples_; }
    int64_t n_samples() const { return (int64_t)samples_.size(); }
protected:
    std::vector<std::string> samples_;
};

// nullptr: the file cannot be opened or is neither VCF nor BCF (the reference: open() == false).
std::unique_ptr<VariantSource> open_variant_source(const std::string &path);

}  // namespace nph
