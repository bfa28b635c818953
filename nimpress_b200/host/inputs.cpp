#include "inputs.hpp"

#include <algorithm>
#include <fstream>

namespace nph {

// Nim readLine: a line ends at \n or \r\n; a final line without newline is still a line.
static bool read_line(std::istream &in, std::string &line) {
    if (!std::getline(in, line)) return false;
    if (!line.empty() && line.back() == '\r') line.pop_back();
    return true;
}

bool ScoreFile::load(const std::string &path) {
    std::ifstream in(path, std::ios::binary);
    if (!in) return false;
    std::string line;
    std::string *hdr[4] = { &name, &desc, &cite, &genomever };
    for (int i = 0; i < 4; i++) {
        if (!read_line(in, line)) throw InputError("score file: truncated header");
        rstrip_nim(line);
        *hdr[i] = line;
    }
    if (!read_line(in, line)) throw InputError("score file: missing offset line");
    rstrip_nim(line);
    offset = parse_float_nim(line, "score file offset");
    entries.clear();
    while (read_line(in, line)) {
        rstrip_nim(line);
        std::vector<std::string> f = split_char(line, '\t');
        if (f.size() != 6) throw InputError("score file: expected 6 tab-separated fields, got " + std::to_string(f.size()));
        ScoreEntry e;
        e.contig = f[0];
        e.pos = parse_int_nim(f[1], "score file pos");
        e.refseq = f[2];
        e.easeq = f[3];
        e.beta = parse_float_nim(f[4], "score file beta");
        e.eaf = parse_float_nim(f[5], "score file eaf");
        entries.push_back(std::move(e));
    }
    return true;
}

bool GenomeIntervals::load(const std::string &path) {
    std::ifstream in(path, std::ios::binary);
    if (!in) return false;
    init = false;
    by_contig.clear();
    std::string line;
    while (read_line(in, line)) {
        rstrip_nim(line);
        std::vector<std::string> f = split_char(line, '\t');
        if (f.size() < 3) throw InputError("coverage BED: expected at least 3 tab-separated fields");
        by_contig[f[0]].emplace_back(parse_int_nim(f[1], "BED start"), parse_int_nim(f[2], "BED end"));
    }
    for (auto &kv : by_contig) std::sort(kv.second.begin(), kv.second.end());
    init = true;
    return true;
}

bool GenomeIntervals::covers(const ScoreEntry &e) const {
    auto it = by_contig.find(e.contig);
    if (it == by_contig.end()) return false;
    const int64_t stop = e.stop();
    for (const auto &iv : it->second) {
        if (iv.first >= e.pos) break;                   // sorted by start: no later interval has start < pos
        if (iv.second >= stop) return true;             // contains (:310-311)
    }
    return false;
}

}  // namespace nph
