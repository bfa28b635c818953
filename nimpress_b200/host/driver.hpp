// driver.hpp -- host side of computePolygenicScores (src/nimpress.nim:592-649) over the CUDA
// C ABI.  The host keeps what the reference does per locus on strings and files -- coverage
// lookup, findVariant, the FILTER test, eaidx -- and hands decode / tally / decision /
// imputation / accumulation to libnimpress_cuda.
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/nimpress_cuda.h"
#include "inputs.hpp"
#include "variant_source.hpp"

namespace nph {

// enums mirror ImputeMethodLocus / Missing / Sample (src/nimpress.nim:412-414)
struct ScoreParams {
    int imp_locus = NPC_LOCUS_PS;          // --imp-locus   [default: ps]
    int imp_missing = NPC_MISSING_HOMREF;  // --imp-missing [default: homref]
    int imp_sample = NPC_SAMPLE_INT_PS;    // --imp-sample  [default: int_ps]
    double maxmis = 0.05;                  // --maxmis
    double afmisp = 0.001;                 // --afmisp
    int64_t mincs = 100;                   // --mincs
    bool ignorefilt = false;               // --ignorefilt
    bool use_cov = false;                  // --cov given (restrictToCoveredRgns)
    int device = 0;                        // CUDA device (not a reference option)
    std::vector<int> devices;              // several GPUs: the score rows are split into contiguous ranges, one per device (empty: `device`)
    bool exact_order = false;              // npc_set_exact_order: bit-for-bit reference summation order
    bool use_ds = false;                   // score FORMAT/DS (fp32 expected ALT dosage) instead of FORMAT/GT -- not a reference option
};

struct ScoreResult {
    std::vector<std::string> samples;
    std::vector<double> scores;            // per sample, after /(2*nloci) and +offset
    std::vector<npc_locus> loci;           // per score row, score-file order
    int64_t nloci = 0;
    std::string warnings;                  // the "WARN ..." lines the reference logs, in its order
    int64_t devices = 1;                   // GPUs that scored (npc_reduce combined their partial sums when > 1)
    int64_t records_read = 0, records_matched = 0, rounds = 0, index_seeks = 0;   // index_seeks > 0: the .tbi / .csi index was used
};

// Host-side part of getImputedDosages that needs no genotypes: coverage (:526), the streaming
// equivalent of findVariant (:353-364: first record in file order overlapping contig:pos-stop with
// REF == refseq and the effect allele equal to REF or present in ALT), eaidx (:375-379) and the
// FILTER test (:553).  Pure CPU: unit-tested without a GPU.
class Matcher {
public:
    Matcher(const ScoreFile &score, const GenomeIntervals &cov, const ScoreParams &p);
    // entries settled by this record (their kind / eaidx / filter text are set); empty if none
    const std::vector<int64_t> &match(const VariantRecord &rec);
    const std::vector<int64_t> &match_last() const { return hits_; }   // what the last match() returned
    void finish();                                  // everything still pending is ABSENT (:536)
    bool has_contig(const std::string &c) const { return index_.count(c) != 0; }
    int64_t n_lookup() const { return n_lookup_; }
    // 1-based inclusive [pos, stop] of every entry that needs a record lookup, per contig (for the index-driven reader)
    void add_spans(std::unordered_map<std::string, std::vector<std::pair<int64_t, int64_t>>> &spans) const;

    static constexpr int32_t PENDING = -1;
    std::vector<int32_t> kind, eaidx;               // per score entry; kind is NPC_KIND_* or PENDING
    std::vector<std::string> filter_text;           // FILTER string of the matched record (FILTER rows)
    std::vector<uint8_t> contig_in_bed;
private:
    struct ContigIndex { std::vector<std::pair<int64_t, int64_t>> by_pos; int64_t max_reflen = 1; };
    const std::vector<ScoreEntry> &E_;
    bool ignorefilt_;
    std::unordered_map<std::string, ContigIndex> index_;
    std::vector<int64_t> hits_;
    int64_t n_lookup_ = 0;
};

// Throws InputError where the reference raises; std::runtime_error on CUDA / library failure.
// false: the genotype file cannot be opened (the reference: FATAL + quit(-1), :728-730).
bool compute_polygenic_scores(const ScoreFile &score, const std::string &genotype_path, const GenomeIntervals &cov,
                              const ScoreParams &p, ScoreResult &out);

// Several score files over ONE pass of the genotype file (BASELINE config 4).  Each result is what
// compute_polygenic_scores gives for that file alone (the reference: one nimpress run per file).
bool compute_polygenic_scores_multi(const std::vector<const ScoreFile *> &scores, const std::string &genotype_path,
                                    const GenomeIntervals &cov, const ScoreParams &p, std::vector<ScoreResult> &outs);

}  // namespace nph
