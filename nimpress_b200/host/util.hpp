// util.hpp -- small string / number helpers that mirror the Nim stdlib calls the reference makes
// (strip(leading=false), split('\t'), parseFloat, parseInt, `$`(float)).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

namespace nph {

// Errors that the reference turns into an exception / failed doAssert (exit code 1).
struct InputError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// Nim strutils.strip(leading = false): trailing whitespace {' ', \t, \v, \r, \n, \f}
inline void rstrip_nim(std::string &s) {
    size_t n = s.size();
    while (n && (s[n - 1] == ' ' || s[n - 1] == '\t' || s[n - 1] == '\v' || s[n - 1] == '\r' || s[n - 1] == '\n' ||
                 s[n - 1] == '\f'))
        n--;
    s.resize(n);
}

// Nim split(sep: char): keeps empty fields
inline std::vector<std::string> split_char(const std::string &s, char sep) {
    std::vector<std::string> out;
    size_t b = 0;
    for (;;) {
        size_t e = s.find(sep, b);
        if (e == std::string::npos) { out.emplace_back(s, b); break; }
        out.emplace_back(s, b, e - b);
        b = e + 1;
    }
    return out;
}

// Nim parseFloat: the whole string, "nan"/"inf" spellings accepted, else ValueError
inline double parse_float_nim(const std::string &s, const char *what) {
    if (s.empty()) throw InputError(std::string("invalid float (empty): ") + what);
    char *end = nullptr;
    double v = std::strtod(s.c_str(), &end);
    if (end == s.c_str() || *end) throw InputError("invalid float: " + s + " (" + what + ")");
    return v;
}

inline int64_t parse_int_nim(const std::string &s, const char *what) {
    if (s.empty()) throw InputError(std::string("invalid integer (empty): ") + what);
    char *end = nullptr;
    long long v = std::strtoll(s.c_str(), &end, 10);
    if (end == s.c_str() || *end) throw InputError("invalid integer: " + s + " (" + what + ")");
    return v;
}

// Nim (< 1.6) `$`(float): C "%.16g", ".0" appended to integral-looking text, nan / inf / -inf.
// This is the format of the reference's output lines (src/nimpress.nim:753) and WARN texts.
inline std::string format_float_nim(double v) {
    if (std::isnan(v)) return "nan";
    if (std::isinf(v)) return v > 0 ? "inf" : "-inf";
    char buf[64];
    int n = std::snprintf(buf, sizeof buf, "%.16g", v);
    bool plain = true;
    for (int i = 0; i < n; i++) {
        if (buf[i] == ',') buf[i] = '.';
        if (!(buf[i] == '-' || (buf[i] >= '0' && buf[i] <= '9'))) plain = false;
    }
    std::string s(buf, n);
    if (plain) s += ".0";
    return s;
}

}  // namespace nph
