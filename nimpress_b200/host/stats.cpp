// stats.cpp -- the statistics behind the reference's AF-mismatch WARN lines.
//
// lbinom / dbinom / betacf / betai / pbinom below (up to `count_tail`) TRANSLITERATE src/nimpress.nim:51-152 into C++,
// operation for operation: the WARN lines print p-value-dependent decisions, so these functions must return the
// reference's doubles bit for bit (the Numerical-Recipes continued fraction included) and there is no second way to
// write them.  The same text exists once more in oracle/nimpress_oracle.c, which is the checker.  What is new here is
// binom_test's evaluation: the reference enumerates all n + 1 outcomes (:155-188); `count_tail` finds the same
// integration limits by bisection on the monotone tails (tested equal to the enumeration, tests/test_host_cpu.py).
#include "stats.hpp"

#include <algorithm>
#include <cmath>

namespace nph {

static double lbinom(int64_t n, int64_t k) {          // :51
    return std::lgamma((double)n + 1.0) - std::lgamma((double)k + 1.0) - std::lgamma((double)(n - k) + 1.0);
}

double dbinom(int64_t x, int64_t n, double p) {
    if ((x == 0 && p == 0.0) || (x == n && p == 1.0)) return 1.0;
    return std::exp(lbinom(n, x) + (double)x * std::log(p) + (double)(n - x) * std::log(1.0 - p));
}

static double betacf(double a, double b, double x) {  // :63-117
    const double qab = a + b, qap = a + 1.0, qam = a - 1.0, FPMIN = 1.0e-30, EPS = 3.0e-7;
    double c = 1.0, d = 1.0 - qab * x / qap;
    if (std::fabs(d) < FPMIN) d = FPMIN;
    d = 1.0 / d;
    double h = d;
    for (int m = 1; m <= 100; m++) {
        const double mf = m;
        double aa = mf * (b - mf) * x / ((qam + 2 * mf) * (a + 2 * mf));
        d = 1.0 + aa * d; if (std::fabs(d) < FPMIN) d = FPMIN;
        c = 1.0 + aa / c; if (std::fabs(c) < FPMIN) c = FPMIN;
        d = 1.0 / d;
        h *= d * c;
        aa = -(a + mf) * (qab + mf) * x / ((a + 2 * mf) * (qap + 2 * mf));
        d = 1.0 + aa * d; if (std::fabs(d) < FPMIN) d = FPMIN;
        c = 1.0 + aa / c; if (std::fabs(c) < FPMIN) c = FPMIN;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (std::fabs(del - 1.0) < EPS) return h;
    }
    return NAN;
}

double betai(double a, double b, double x) {
    if (!(x >= 0.0 && x <= 1.0)) return NAN;
    if (a == 0.0 || b == 0.0) return INFINITY;
    if (x == 0.0) return 0.0;
    if (x == 1.0) return 1.0;
    const double bt = std::exp(std::lgamma(a + b) - std::lgamma(a) - std::lgamma(b) + a * std::log(x) + b * std::log(1.0 - x));
    if (x < (a + 1.0) / (a + b + 2.0)) return bt * betacf(a, b, x) / a;
    return 1.0 - bt * betacf(b, a, 1.0 - x) / b;
}

double pbinom(int64_t x, int64_t n, double p) {
    if (x < 0) return 0.0;
    if (x == n) return 1.0;
    return 1.0 - betai((double)x + 1.0, (double)(n - x), p);
}

// number of xi in [lo, hi] with dbinom(xi) <= thr, where dbinom is non-increasing (dir > 0) or
// non-decreasing (dir < 0) along the range up to rounding noise: bisect for the crossing, then
// settle the +-W neighbourhood by direct evaluation so that noise near the crossing is counted
// exactly as the reference's enumeration would count it.
static int64_t count_tail(int64_t lo, int64_t hi, int64_t n, double p, double thr, int dir) {
    if (lo > hi) return 0;
    const int64_t W = 8;
    auto below = [&](int64_t xi) { return dbinom(xi, n, p) <= thr; };
    int64_t a = lo, b = hi;                       // invariant for dir>0: all xi > b are below; find first below
    if (dir > 0) {
        if (!below(hi)) return 0;
        while (a < b) { int64_t m = a + (b - a) / 2; if (below(m)) b = m; else a = m + 1; }
        int64_t first = a, cnt = hi - first + 1;  // [first, hi] assumed below
        for (int64_t xi = std::max(lo, first - W); xi < first; xi++) if (below(xi)) cnt++;
        for (int64_t xi = first; xi <= std::min(hi, first + W); xi++) if (!below(xi)) cnt--;
        return cnt;
    } else {
        if (!below(lo)) return 0;
        while (a < b) { int64_t m = a + (b - a + 1) / 2; if (below(m)) a = m; else b = m - 1; }
        int64_t last = a, cnt = last - lo + 1;    // [lo, last] assumed below
        for (int64_t xi = last + 1; xi <= std::min(hi, last + W); xi++) if (below(xi)) cnt++;
        for (int64_t xi = std::max(lo, last - W); xi <= last; xi++) if (!below(xi)) cnt--;
        return cnt;
    }
}

double binom_test(int64_t x, int64_t n, double p) {
    if (p == 0.0) return x == 0 ? 1.0 : 0.0;
    if (p == 1.0) return x == n ? 1.0 : 0.0;
    const double probx = dbinom(x, n, p), expected = (double)n * p;
    if (std::fabs((double)x / expected - 1.0) < 1.0e-6) return 1.0;
    const double thr = probx * (1.0 + 1.0e-7);
    if ((double)x < expected) {
        // :176-181  xi in ceil(E)..n, beyond the mode: dbinom non-increasing
        const int64_t y = count_tail((int64_t)std::ceil(expected), n, n, p, thr, +1);
        return pbinom(x, n, p) + (1.0 - pbinom(n - y, n, p));
    }
    // :182-187  xi in 0..floor(E), before the mode: dbinom non-decreasing
    const int64_t y = count_tail(0, (int64_t)std::floor(expected), n, p, thr, -1);
    return pbinom(y - 1, n, p) + (1.0 - pbinom(x - 1, n, p));
}

}  // namespace nph
