// util.hpp -- small string / number helpers that mirror the Nim stdlib calls the reference makes
// (strip(leading=false), split('\t'), parseFloat, parseInt, `$`(float)).
#pragma once
#include <cctype>
#include <charconv>
#include <cstring>
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
    // strtod takes more than Nim's parseFloat does: leading white space, hex floats (0x1p3), nan(chars)
    {
        const unsigned char c0 = (unsigned char)s[0];
        bool bad = isspace(c0) != 0;
        for (size_t i = 0; i + 1 < s.size() && !bad; i++)
            if ((s[i] == '0' && (s[i + 1] == 'x' || s[i + 1] == 'X')) || s[i] == '(') bad = true;
        if (bad) throw InputError("invalid float: " + s + " (" + what + ")");
    }
    char *end = nullptr;
    double v = std::strtod(s.c_str(), &end);
    if (end == s.c_str() || *end) throw InputError("invalid float: " + s + " (" + what + ")");
    return v;
}

inline int64_t parse_int_nim(const std::string &s, const char *what) {
    if (s.empty() || isspace((unsigned char)s[0])) throw InputError(std::string("invalid integer: '") + s + "' (" + what + ")");
    char *end = nullptr;
    long long v = std::strtoll(s.c_str(), &end, 10);
    if (end == s.c_str() || *end) throw InputError("invalid integer: " + s + " (" + what + ")");
    return v;
}

// Nim (< 1.6) `$`(float): C "%.16g", ".0" appended to integral-looking text, nan / inf / -inf.
// This is the format of the reference's output lines (src/nimpress.nim:753) and WARN texts.
// Appends to buf (>= 40 bytes free); returns the length.  std::to_chars(general, 16) is specified to
// give printf("%.16g") in the C locale, several times faster.
inline int format_float_nim_to(double v, char *buf) {
    if (std::isnan(v)) { memcpy(buf, "nan", 3); return 3; }
    if (std::isinf(v)) { if (v > 0) { memcpy(buf, "inf", 3); return 3; } memcpy(buf, "-inf", 4); return 4; }
    char *end = std::to_chars(buf, buf + 32, v, std::chars_format::general, 16).ptr;
    int n = (int)(end - buf);
    bool plain = true;
    for (int i = 0; i < n; i++) if (!(buf[i] == '-' || (buf[i] >= '0' && buf[i] <= '9'))) { plain = false; break; }
    if (plain) { buf[n++] = '.'; buf[n++] = '0'; }
    return n;
}
inline std::string format_float_nim(double v) {
    char buf[48];
    return std::string(buf, format_float_nim_to(v, buf));
}

}  // namespace nph
