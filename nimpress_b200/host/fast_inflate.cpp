#include "fast_inflate.hpp"

#include <cstring>

namespace nph {

namespace {

// table entry: bits 0-3 code bits to consume at this level | 4-5 kind | 6-9 extra-bit count (links: subtable index bits)
//              | 10-31 payload (literal byte, base length / distance, or subtable offset)
enum { K_LIT = 0, K_BASE = 1, K_EOB = 2, K_LINK = 3 };
constexpr uint32_t INVALID = (uint32_t)K_LINK << 4;              // a link of zero bits: never built, marks unassigned codes
inline uint32_t mk(uint32_t bits, uint32_t kind, uint32_t extra, uint32_t payload) { return bits | (kind << 4) | (extra << 6) | (payload << 10); }
inline uint32_t e_bits(uint32_t e) { return e & 15u; }
inline uint32_t e_kind(uint32_t e) { return (e >> 4) & 3u; }
inline uint32_t e_extra(uint32_t e) { return (e >> 6) & 15u; }
inline uint32_t e_payload(uint32_t e) { return e >> 10; }

constexpr int LIT_P = 10, DIST_P = 8;
const uint16_t LEN_BASE[29] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258 };
const uint8_t LEN_EXTRA[29] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };
const uint16_t DIST_BASE[30] = { 1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289,
                                 16385, 24577 };
const uint8_t DIST_EXTRA[30] = { 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13 };

inline uint32_t symbol_entry(bool dist_table, int sym, uint32_t bits) {
    if (dist_table) return sym < 30 ? mk(bits, K_BASE, DIST_EXTRA[sym], DIST_BASE[sym]) : INVALID;
    if (sym < 256) return mk(bits, K_LIT, 0, (uint32_t)sym);
    if (sym == 256) return mk(bits, K_EOB, 0, 0);
    return sym < 286 ? mk(bits, K_BASE, LEN_EXTRA[sym - 257], LEN_BASE[sym - 257]) : INVALID;
}

// canonical Huffman code -> primary table of 2^P entries indexed by the next P input bits (LSB first), codes longer
// than P through subtables of 2^(maxlen - P) entries.  false: over-subscribed lengths or table overflow.
bool build(const uint8_t *lens, int n, int P, bool dist_table, uint32_t *table, int cap) {
    int count[16] = { 0 };
    for (int i = 0; i < n; i++) count[lens[i]]++;
    count[0] = 0;
    int left = 1, maxlen = 0;
    for (int l = 1; l <= 15; l++) {
        left = left * 2 - count[l];
        if (left < 0) return false;
        if (count[l]) maxlen = l;
    }
    uint32_t next[16], code = 0;
    for (int l = 1; l <= 15; l++) { code = (code + (uint32_t)count[l - 1]) << 1; next[l] = code; }
    const int psize = 1 << P, sub_bits = maxlen > P ? maxlen - P : 0;
    for (int i = 0; i < psize; i++) table[i] = INVALID;
    int used = psize;
    for (int sym = 0; sym < n; sym++) {
        const int l = lens[sym];
        if (!l) continue;
        uint32_t c = next[l]++, rev = 0;
        for (int i = 0; i < l; i++) { rev = (rev << 1) | (c & 1u); c >>= 1; }
        if (l <= P) {
            const uint32_t e = symbol_entry(dist_table, sym, (uint32_t)l);
            for (uint32_t i = rev; i < (uint32_t)psize; i += 1u << l) table[i] = e;
        } else {
            const uint32_t prefix = rev & (uint32_t)(psize - 1);
            if (table[prefix] == INVALID) {
                if (used + (1 << sub_bits) > cap) return false;
                for (int i = 0; i < (1 << sub_bits); i++) table[used + i] = INVALID;
                table[prefix] = mk((uint32_t)P, K_LINK, (uint32_t)sub_bits, (uint32_t)used);
                used += 1 << sub_bits;
            }
            const uint32_t base = e_payload(table[prefix]), e = symbol_entry(dist_table, sym, (uint32_t)(l - P));
            for (uint32_t i = rev >> P; i < (1u << sub_bits); i += 1u << (l - P)) table[base + i] = e;
        }
    }
    return true;
}

inline uint64_t load64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
inline void store64(uint8_t *p, uint64_t v) { memcpy(p, &v, 8); }

}  // namespace

bool fast_inflate(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len, FastInflateTables &t) {
    const uint8_t *ip = in, *const in_end = in + in_len;
    uint8_t *op = out, *const out_end = out + out_len;
    uint64_t bb = 0;
    uint32_t bc = 0;
    // the bit buffer runs up to 8 bytes ahead of what has been consumed, so ip may reach in_end + 8 on a good stream
    // (the caller guarantees 16 readable bytes past in_len); further than that the stream is corrupt
#define REFILL()                                          \
    do {                                                  \
        if (ip > in_end + 8) return false;                \
        bb |= load64(ip) << bc;                           \
        ip += (63 - bc) >> 3;                             \
        bc |= 56;                                         \
    } while (0)
#define TAKE(nb) (bb >>= (nb), bc -= (nb))

    for (;;) {
        REFILL();
        const uint32_t final_block = (uint32_t)bb & 1u, type = ((uint32_t)bb >> 1) & 3u;
        TAKE(3);
        if (type == 0) {                                                // stored: to the byte boundary, LEN, ~LEN, bytes
            TAKE(bc & 7u);
            ip -= bc >> 3;                                              // whole bytes still in the bit buffer go back
            bb = 0; bc = 0;
            if (in_end - ip < 4) return false;
            const uint32_t len = ip[0] | (ip[1] << 8), nlen = ip[2] | (ip[3] << 8);
            ip += 4;
            if ((len ^ nlen) != 0xFFFFu || (size_t)(in_end - ip) < len || (size_t)(out_end - op) < len) return false;
            memcpy(op, ip, len);
            op += len; ip += len;
        } else if (type == 3) {
            return false;
        } else {
            if (type == 1) {                                            // fixed code
                uint8_t lens[288];
                for (int i = 0; i < 144; i++) lens[i] = 8;
                for (int i = 144; i < 256; i++) lens[i] = 9;
                for (int i = 256; i < 280; i++) lens[i] = 7;
                for (int i = 280; i < 288; i++) lens[i] = 8;
                if (!build(lens, 288, LIT_P, false, t.lit, (int)(sizeof t.lit / 4))) return false;
                for (int i = 0; i < 32; i++) lens[i] = 5;
                if (!build(lens, 32, DIST_P, true, t.dist, (int)(sizeof t.dist / 4))) return false;
            } else {                                                    // dynamic code: the code lengths are themselves Huffman coded
                const uint32_t hlit = ((uint32_t)bb & 31u) + 257, hdist = (((uint32_t)bb >> 5) & 31u) + 1, hclen = (((uint32_t)bb >> 10) & 15u) + 4;
                TAKE(14);
                if (hlit > 286 || hdist > 30) return false;
                static const uint8_t order[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
                uint8_t cl[19] = { 0 };
                REFILL();
                for (uint32_t i = 0; i < hclen; i++) {
                    if (bc < 3) REFILL();
                    cl[order[i]] = (uint8_t)(bb & 7u);
                    TAKE(3);
                }
                uint32_t pre[128];                                      // 7-bit primary table, no subtables (lengths <= 7)
                {
                    uint32_t big[128 + 8];
                    if (!build(cl, 19, 7, false, big, 128)) return false;
                    memcpy(pre, big, sizeof pre);
                }
                uint8_t lens[288 + 32] = { 0 };
                for (uint32_t i = 0; i < hlit + hdist;) {
                    REFILL();
                    const uint32_t e = pre[bb & 127u];
                    if (e == INVALID) return false;
                    TAKE(e_bits(e));
                    const uint32_t sym = e_payload(e);                  // symbols 0..18 come out as K_LIT entries
                    if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
                    uint32_t rep, val = 0;
                    if (sym == 16) { if (i == 0) return false; val = lens[i - 1]; rep = 3 + ((uint32_t)bb & 3u); TAKE(2); }
                    else if (sym == 17) { rep = 3 + ((uint32_t)bb & 7u); TAKE(3); }
                    else { rep = 11 + ((uint32_t)bb & 127u); TAKE(7); }
                    if (i + rep > hlit + hdist) return false;
                    while (rep--) lens[i++] = (uint8_t)val;
                }
                if (lens[256] == 0) return false;                       // no end-of-block code
                if (!build(lens, (int)hlit, LIT_P, false, t.lit, (int)(sizeof t.lit / 4))) return false;
                if (!build(lens + hlit, (int)hdist, DIST_P, true, t.dist, (int)(sizeof t.dist / 4))) return false;
            }
            // ---- symbols ----
            // >= 56 bits after a refill: a literal/length code (15 + 5) and a distance code (15 + 13) fit.  The entry of the
            // NEXT symbol is looked up before a match is copied, so that the table load overlaps the copy.
            REFILL();
            uint32_t e = t.lit[bb & ((1u << LIT_P) - 1u)];
            for (;;) {
                if (e_kind(e) == K_LINK) {
                    if (e == INVALID) return false;
                    TAKE(LIT_P);
                    e = t.lit[e_payload(e) + (bb & ((1u << e_extra(e)) - 1u))];
                    if (e_kind(e) == K_LINK) return false;
                }
                TAKE(e_bits(e));
                if (e_kind(e) == K_LIT) {
                    if (op >= out_end) return false;
                    *op++ = (uint8_t)e_payload(e);
                    // more literals from the same refill while they are there (short runs of literals are common):
                    // 56 bits hold at least three further primary-table codes
                    for (int k = 0; k < 3; k++) {
                        const uint32_t e2 = t.lit[bb & ((1u << LIT_P) - 1u)];
                        if (e_kind(e2) != K_LIT || op >= out_end) break;
                        TAKE(e_bits(e2));
                        *op++ = (uint8_t)e_payload(e2);
                    }
                    REFILL();
                    e = t.lit[bb & ((1u << LIT_P) - 1u)];
                    continue;
                }
                if (e_kind(e) == K_EOB) break;
                uint32_t len = e_payload(e) + ((uint32_t)bb & ((1u << e_extra(e)) - 1u));
                TAKE(e_extra(e));
                uint32_t d = t.dist[bb & ((1u << DIST_P) - 1u)];
                if (e_kind(d) == K_LINK) {
                    if (d == INVALID) return false;
                    TAKE(DIST_P);
                    d = t.dist[e_payload(d) + (bb & ((1u << e_extra(d)) - 1u))];
                    if (e_kind(d) == K_LINK) return false;
                }
                if (e_kind(d) != K_BASE) return false;
                TAKE(e_bits(d));
                const uint32_t dist = e_payload(d) + ((uint32_t)bb & ((1u << e_extra(d)) - 1u));
                TAKE(e_extra(d));
                if (dist > (size_t)(op - out) || len > (size_t)(out_end - op)) return false;
                REFILL();
                e = t.lit[bb & ((1u << LIT_P) - 1u)];
                const uint8_t *src = op - dist;
                uint8_t *const stop = op + len;
                if (dist >= 8) {                                        // word copies; they may run up to 13 bytes past `stop` (caller's slack: 16)
                    store64(op, load64(src)); store64(op + 8, load64(src + 8));        // most matches are short: no loop, no loop-exit misprediction
                    if (len > 16) { op += 16; src += 16; do { store64(op, load64(src)); op += 8; src += 8; } while (op < stop); }
                } else if (dist == 1) {                                 // a run of one byte
                    const uint64_t v = 0x0101010101010101ull * *src;
                    do { store64(op, v); op += 8; } while (op < stop);
                } else if (dist == 2 || dist == 4) {                    // period 2 or 4: one replicated word (GT pairs repeat at distance 2)
                    uint64_t v;
                    if (dist == 2) { uint16_t h; memcpy(&h, src, 2); v = 0x0001000100010001ull * h; }
                    else { uint32_t w; memcpy(&w, src, 4); v = 0x0000000100000001ull * w; }
                    do { store64(op, v); op += 8; } while (op < stop);
                } else {
                    // period < 8: copy bytes until the gap back to the source is a multiple of the period >= 8, then words
                    const uint32_t m = ((8 + dist - 1) / dist) * dist;
                    uint32_t k = len < m ? len : m;
                    while (k--) *op++ = *src++;
                    src = op - m;
                    while (op < stop) { store64(op, load64(src)); op += 8; src += 8; }
                }
                op = stop;
            }
        }
        if (final_block) break;
    }
#undef REFILL
#undef TAKE
    // bytes actually consumed must lie inside the input
    if ((size_t)(ip - in) - (bc >> 3) > in_len) return false;
    return op == out_end;
}

}  // namespace nph
