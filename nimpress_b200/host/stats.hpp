// stats.hpp -- the exact two-sided binomial test behind nimpress's allele-frequency warnings
// (src/nimpress.nim:50-188).  Same formulas as the reference (lgamma-based dbinom, NRC
// incomplete beta); the only change is HOW the integration limits are found: the reference
// enumerates dbinom over half of 0..n (O(n) per locus -- at biobank scale it spends most of its
// run time here), this finds the same limit by bisection on the monotone
// tail and then re-checks the neighbourhood of the crossing by direct evaluation (O(log n)).
#pragma once
#include <cstdint>

namespace nph {
double dbinom(int64_t x, int64_t n, double p);        // :54-60
double betai(double a, double b, double x);           // :120-134
double pbinom(int64_t x, int64_t n, double p);        // :138-152
double binom_test(int64_t x, int64_t n, double p);    // :155-188
}  // namespace nph
