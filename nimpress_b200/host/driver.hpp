// driver.hpp -- host side of computePolygenicScores (src/nimpress.nim:592-649) over the CUDA
// C ABI.  The host keeps what the reference does per locus on strings and files -- coverage
// lookup, findVariant, the FILTER test, eaidx -- and hands decode / tally / decision /
// imputation / accumulation to libnimpress_cuda.
#pragma once
#include <cstdint>
#include <string>
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
};

struct ScoreResult {
    std::vector<std::string> samples;
    std::vector<double> scores;            // per sample, after /(2*nloci) and +offset
    std::vector<npc_locus> loci;           // per score row, score-file order
    int64_t nloci = 0;
    std::string warnings;                  // the "WARN ..." lines the reference logs, in its order
    int64_t records_read = 0, records_matched = 0, rounds = 0;
};

// Throws InputError where the reference raises; std::runtime_error on CUDA / library failure.
void compute_polygenic_scores(const ScoreFile &score, VariantSource &vcf, const GenomeIntervals &cov,
                              const ScoreParams &p, ScoreResult &out);

}  // namespace nph
