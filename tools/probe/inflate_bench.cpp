// Single-thread rate of the BGZF block decoder (nimpress_b200/host/fast_inflate.cpp) against zlib's inflate on GT-like data.
//   g++ -O2 -std=c++17 -o /tmp/inflate_bench tools/probe/inflate_bench.cpp nimpress_b200/host/fast_inflate.cpp -lz && /tmp/inflate_bench
#include <zlib.h>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>
#include "../../nimpress_b200/host/fast_inflate.hpp"

int main(int argc, char **argv) {
    const int level = argc > 1 ? atoi(argv[1]) : 1;
    const size_t BLOCK = 0xFF00, NB = 2000;
    std::mt19937_64 rng(1);
    std::vector<std::vector<uint8_t>> raw(NB), comp(NB);
    size_t raw_total = 0, comp_total = 0;
    double af = 0.2;
    for (size_t b = 0; b < NB; b++) {
        raw[b].resize(BLOCK + 16);
        if (b % 16 == 0) af = std::uniform_real_distribution<double>(0.01, 0.5)(rng);      // a new variant every MB
        for (size_t i = 0; i < BLOCK; i++) {
            const double u = std::uniform_real_distribution<double>(0, 1)(rng);
            raw[b][i] = u < 0.005 ? 0 : (u < af ? 4 : 2);                                 // BCF GT bytes: (allele + 1) << 1
        }
        comp[b].resize(BLOCK + 1024);
        z_stream z; memset(&z, 0, sizeof z);
        deflateInit2(&z, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
        z.next_in = raw[b].data(); z.avail_in = BLOCK; z.next_out = comp[b].data(); z.avail_out = comp[b].size() - 16;
        deflate(&z, Z_FINISH);
        comp[b].resize(z.total_out + 16);
        deflateEnd(&z);
        raw_total += BLOCK; comp_total += z.total_out;
    }
    printf("level %d: %zu blocks, ratio %.2f\n", level, NB, (double)raw_total / comp_total);
    std::vector<uint8_t> out(BLOCK + 16);
    nph::FastInflateTables t;
    for (int rep = 0; rep < 2; rep++) {
        auto t0 = std::chrono::steady_clock::now();
        size_t bad = 0;
        for (size_t b = 0; b < NB; b++) {
            if (!nph::fast_inflate(comp[b].data(), comp[b].size() - 16, out.data(), BLOCK, t) || memcmp(out.data(), raw[b].data(), BLOCK)) bad++;
        }
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("fast_inflate: %.3f GB/s out (%zu bad)\n", raw_total / dt / 1e9, bad);
        t0 = std::chrono::steady_clock::now();
        for (size_t b = 0; b < NB; b++) {
            z_stream z; memset(&z, 0, sizeof z);
            inflateInit2(&z, -15);
            z.next_in = comp[b].data(); z.avail_in = comp[b].size() - 16; z.next_out = out.data(); z.avail_out = BLOCK;
            inflate(&z, Z_FINISH);
            inflateEnd(&z);
        }
        dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("zlib inflate: %.3f GB/s out\n", raw_total / dt / 1e9);
    }
}
