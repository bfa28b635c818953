// inputs.hpp -- the two small text inputs of nimpress: the polygenic score definition
// (src/nimpress.nim:195-254) and the coverage BED (src/nimpress.nim:262-345).
#pragma once
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "util.hpp"

namespace nph {

// ScoreEntry (src/nimpress.nim:221-228)
struct ScoreEntry {
    std::string contig, refseq, easeq;
    int64_t pos = 0;
    double beta = 0, eaf = 0;
    int64_t stop() const { return pos + (int64_t)refseq.size() - 1; }   // :230-231
    bool ref_is_ea() const { return refseq == easeq; }
};

// ScoreFile (src/nimpress.nim:195-219): five header lines, then 6-column TSV rows.
struct ScoreFile {
    std::string name, desc, cite, genomever;
    double offset = 0;
    std::vector<ScoreEntry> entries;     // file order = processing order

    // open (:233-244) + items (:247-254), read eagerly.  false: file cannot be opened.
    // Malformed content throws InputError where the reference raises / asserts.
    bool load(const std::string &path);
};

// GenomeIntervals (src/nimpress.nim:262-308).  The reference indexes each contig with lapper
// and then keeps only intervals that `contain` the entry (:310-311); containment implies the
// overlap lapper pre-filters on, so a per-contig scan over intervals sorted by start is
// equivalent (and stops early at start >= pos).
struct GenomeIntervals {
    bool init = false;
    std::unordered_map<std::string, std::vector<std::pair<int64_t, int64_t>>> by_contig;   // (start0, end1)

    bool load(const std::string &path);                                // loadBedIntervals (:278-308)
    bool has_contig(const std::string &c) const { return by_contig.count(c) != 0; }
    bool covers(const ScoreEntry &e) const;                            // isVariantCovered (:313-345), sans logging
};

}  // namespace nph
