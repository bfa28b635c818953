// npc_fused4.cuh -- the default roofline kernel: the fused persistent design of npc_fused.cuh
// (TMA raw ring -> count -> grid-wide tally dependency -> decide -> accumulate, every genotype
// byte read from HBM once) restructured around TILES OF FOUR SCORE ROWS so that the per-genotype
// instruction and shared-memory cost drops by about 40%:
//
//  * COUNT.  A sample's two raw GT bytes (both < 8 on the fast path: alleles REF, ALT1, ALT2 or
//    missing) are folded into one byte 2*(b0 | b1<<3) with two integer ops per word, and a PRMT
//    composes that byte with the (effect allele, row-in-tile) table number into a complete
//    shared-memory address: one LDS.U16 returns BOTH the sample's 2-bit dosage code
//    {0,1,2,3=missing}, already shifted to the row's position (high byte), AND its tally
//    contribution dosage | missing<<4 (low byte).  OR-ing the four rows gives ONE code byte per
//    sample per tile; adding the entries of 4 samples gives the row's tallies -- no second table.
//  * The index ring therefore holds 1 byte per sample per 4 rows (1/8 of the raw data): the
//    grid-wide dependency (see npc_fused.cuh) can lag by dozens of rows at no cost.
//  * DECIDE builds, per tile, two 16-entry fp64 tables T01[b0|b1<<2] = v0[b0]+v1[b1] and
//    T23[b2|b3<<2] = v2[b2]+v3[b3] of the rows' contributions (constant rows and dropped rows fold
//    in as constants).  16 doubles span exactly the 32 banks: every lookup is conflict-free (one
//    256-entry table was measured at 3-way conflicts: its hot entries share 9 of 16 bank pairs).
//  * ACCUMULATE: sums[s] = (sums[s] + T01[..]) + T23[..] -- two 8-byte table loads and two DADDs
//    per sample per FOUR genotypes, no per-row branches.
//
// Rounding: the contributions of rows (0,1) and of rows (2,3) of a tile are added to each other
// first and then, in that order, to the running sum.  Every addend is still the reference's rounded product fl(dosage*beta); only the
// association differs from the reference's left-to-right chain (src/nimpress.nim:639-640), so
// scores agree to a few ulp of the running sum (tests assert <= 1e-12 relative; the contract is
// 1e-9).  The result is deterministic and independent of the launch shape.
//
// EXACT = true (npc_set_exact_order) keeps everything above except the pre-summing: the tile's
// block holds the four rows' 4-entry tables and ACCUMULATE does one lookup and one DADD per
// genotype, rows in order -- the same rounded products and the same rounded adds as
// `scores[i] += dosages[i]*beta` (src/nimpress.nim:640), bit for bit.  It runs on a 1-D grid (every
// CTA sees every row) so that the per-sample chain is the reference's.
#pragma once
#include "npc_fused.cuh"

namespace npc {

constexpr int F4_R = 4;                                  // rows per tile
constexpr uint32_t F4_CODE_TABLES = FUSED_CNT_TABLES * F4_R;   // (T, row-in-tile) -> 64 two-byte entries in a 256-byte slot

struct Fused4Smem {
    uint32_t code, vrow, bars, cntacc, cisgt, risgt, reaidx, vtab, idx, data, total;
    __host__ __device__ static Fused4Smem make(int Sr, int Sc, int slab_stride) {
        Fused4Smem m;
        uint32_t o = 0;
        m.code = o;   o += F4_CODE_TABLES * 256u;                        // first: 256-byte aligned
        m.vtab = o;   o += (uint32_t)Sc * 256u;                          // 256-byte aligned: T01 at +0, T23 at +128
        m.vrow = o;   o += 6u * 128u * 8u;                               // per decider warp: 32 rows x 4 values
        m.bars = o;   o += (2u * Sr + 2u * Sc) * 8u;            o = (o + 127u) & ~127u;
        m.cntacc = o; o += (uint32_t)Sc * F4_R * 16u * 4u;      o = (o + 127u) & ~127u;
        m.cisgt = o;  o += (uint32_t)Sc * 4u;                   o = (o + 127u) & ~127u;
        m.risgt = o;  o += (uint32_t)Sr * 4u;                   o = (o + 127u) & ~127u;
        m.reaidx = o; o += (uint32_t)Sr * F4_R * 4u;            o = (o + 127u) & ~127u;
        m.idx = o;    o += (uint32_t)Sc * (uint32_t)(slab_stride / 2);  o = (o + 127u) & ~127u;
        m.data = o;   o += (uint32_t)Sr * F4_R * (uint32_t)slab_stride;
        m.total = o;
        return m;
    }
};

// dosage code of one int8 diploid sample whose bytes are both < 16: raw nibbles n0, n1
__device__ __forceinline__ int f4_code_from_nibbles(int n0, int n1, int T) {
    const int c0 = n0 >> 1, c1 = n1 >> 1;                // allele+1, 0 = missing; the phase bit is dropped
    if (c0 == 0 || c1 == 0) return 3;
    return (c0 == T) + (c1 == T);
}
// exact decode (any bytes) -> dosage code
__device__ __forceinline__ uint32_t f4_slow_code(uint32_t h, int eaidx) {
    int8_t a[2] = { (int8_t)(h & 0xFF), (int8_t)((h >> 8) & 0xFF) };
    int d; bool miss;
    decode_sample<int8_t>(a, 2, eaidx, d, miss);
    return miss ? 3u : (uint32_t)d;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// (w << 1) + (w >> 4): byte 0 = 2*(b0 | b1<<3), byte 2 = 2*(b2 | b3<<3) when every byte < 8 -- the
// byte offset of the sample's two-byte table entry
__device__ __forceinline__ uint32_t fold_offsets(uint32_t w) {
    uint32_t t, f;                                   // both on the FMA pipe (the ALU pipe is the busy one)
    asm("mad.hi.u32 %0, %1, 0x10000000, 0;" : "=r"(t) : "r"(w));
    asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(f) : "r"(w), "r"(t));
    return f;
}

template <int K, bool EXACT>
__global__ void __launch_bounds__(640, 1)         // <= 16 consumer warps + producer + publisher + <= 2 deciders
k_fused_tile4(const FusedParams P) {
    constexpr int R = F4_R;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int Sr = P.Sr, Sc = P.Sc, L = P.L, NC = P.nc, A = P.A;
    const Fused4Smem M = Fused4Smem::make(Sr, Sc, P.slab_stride);
    const uint32_t sb = smem_u32(smem);
    const uint32_t bar_full = sb + M.bars, bar_rempty = bar_full + 8u * Sr, bar_cnt = bar_rempty + 8u * Sr, bar_lut = bar_cnt + 8u * Sc;
    uint32_t *s_cntacc = reinterpret_cast<uint32_t *>(smem + M.cntacc);
    uint32_t *s_risgt = reinterpret_cast<uint32_t *>(smem + M.risgt);
    uint32_t *s_cisgt = reinterpret_cast<uint32_t *>(smem + M.cisgt);
    int32_t *s_reaidx = reinterpret_cast<int32_t *>(smem + M.reaidx);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // grid = Gs sample slabs x Gr row groups: this CTA owns chunk range `slab` of the sample axis and
    // the contiguous tile range [tile_lo, tile_lo + n_tiles) of the rows
    const int Gs = P.Gs, slab_id = (int)blockIdx.x % Gs, grp = (int)blockIdx.x / Gs;
    const int64_t all_tiles = (P.n_rows + R - 1) / R;
    const int64_t tile_lo = all_tiles * grp / P.Gr, n_tiles = all_tiles * (grp + 1) / P.Gr - tile_lo;
    const int64_t row_lo = tile_lo * R;

    const int64_t C = (P.n + 7) >> 3;
    const int64_t q = C / Gs, rem = C % Gs;
    const int64_t c0 = (int64_t)slab_id * q + min((int64_t)slab_id, rem);
    const int nch = (int)(q + ((int64_t)slab_id < rem ? 1 : 0));
    const uint32_t slab_bytes = (uint32_t)nch * 16u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < Sr; s++) { mbar_init(bar_full + 8u * s, 1); mbar_init(bar_rempty + 8u * s, NC); }
        for (int s = 0; s < Sc; s++) { mbar_init(bar_cnt + 8u * s, NC); mbar_init(bar_lut + 8u * s, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // tables: [T-1][row r][b0 | b1<<3] (two bytes each) = dosage code << (8 + 2r) | dosage | missing << 4
    for (uint32_t i = threadIdx.x; i < F4_CODE_TABLES * 64u; i += blockDim.x) {
        const int idx = i & 63, r = (i >> 6) & 3, T = (int)(i >> 8) + 1;
        const int c = f4_code_from_nibbles(idx & 7, idx >> 3, T);
        reinterpret_cast<uint16_t *>(smem + M.code)[(i >> 6) * 128 + idx] = (uint16_t)((c << (8 + 2 * r)) | (c == 3 ? 16 : c));
    }
    for (uint32_t i = threadIdx.x; i < (uint32_t)Sr * R * (uint32_t)P.slab_stride / 16u; i += blockDim.x)
        reinterpret_cast<uint4 *>(smem + M.data)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    if (warp == NC) {
        // ================= producer ============================================================
        // Row metadata is read 32 rows (8 tiles) at a time, one row per lane, and the NEXT group is
        // already in flight while this one is issued: a dependent global load per tile would cap the
        // producer at one tile per memory round trip, far below the ring's appetite.
        const uint64_t pol = l2_evict_first_policy();
        constexpr int G = 32 / R;                                    // tiles per metadata group
        int s = 0; uint32_t ph = 0;
        const int64_t row_hi = min(P.n_rows, (tile_lo + n_tiles) * R);      // this group's rows: [row_lo, row_hi)
        npc_row nxt;
        {
            const int64_t r = row_lo + lane;
            if (r < row_hi) nxt = P.rows[r];
        }
        for (int64_t t0 = 0; t0 < n_tiles; t0 += G) {
            const npc_row cur = nxt;
            const int64_t my_row = row_lo + t0 * R + lane;
            const bool have = my_row < row_hi;
            {
                const int64_t r = my_row + 32;
                if (r < row_hi) nxt = P.rows[r];
            }
            const bool is_gt = have && cur.kind == NPC_KIND_GT && cur.gt_row >= 0;
            const uint32_t gt_all = __ballot_sync(0xffffffffu, is_gt);
            const int ng = (int)min((int64_t)G, n_tiles - t0);
            for (int j = 0; j < ng; j++) {
                mbar_wait_sleep(bar_rempty + 8u * s, ph ^ 1u, P.aux_sleep_ns);
                const uint32_t gt_mask = (gt_all >> (R * j)) & ((1u << R) - 1u);
                const bool mine = (lane / R) == j;                   // lanes R*j .. R*j+R-1 own this tile's rows
                if (mine) s_reaidx[s * R + (lane % R)] = is_gt ? cur.eaidx : 0;
                if (lane == 0) s_risgt[s] = gt_mask;
                __syncwarp();
                if (lane == 0) mbar_arrive_expect_tx(bar_full + 8u * s, (uint32_t)__popc(gt_mask) * slab_bytes);
                __syncwarp();
                if (mine && is_gt)
                    tma_load_1d(sb + M.data + (uint32_t)(s * R + (lane % R)) * (uint32_t)P.slab_stride,
                                P.gt + (int64_t)cur.gt_row * P.row_stride + c0 * 16, slab_bytes, bar_full + 8u * s, pol);
                if (++s == Sr) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == NC + 1) {
        // ================= publisher ===========================================================
        int s = 0; uint32_t ph = 0;
        for (int64_t t = 0; t < n_tiles; t++) {
            mbar_wait_sleep(bar_cnt + 8u * s, ph, P.aux_sleep_ns);
            const int nr = (int)min((int64_t)R, P.n_rows - (tile_lo + t) * R);
            if (lane < nr) {
                ull miss = 0, eff = 0;
                if ((s_cisgt[s] >> lane) & 1u)
                    for (int w = 0; w < NC; w++) {
                        const uint32_t v = s_cntacc[(s * R + lane) * 16 + w];
                        miss += v >> 16; eff += v & 0xFFFFu;
                    }
                red_relaxed_gpu_add_u64(P.counts + (tile_lo + t) * R + lane, (1ull << 56) | (miss << FUSED_CNT_BITS) | eff);
            }
            if (++s == Sc) { s = 0; ph ^= 1u; }
        }
    } else if (warp > NC + 1) {
        // ================= deciders ============================================================
        // One pass decides a GROUP of 8 tiles = 32 rows, one row per lane: the global round trip of
        // the tally word (microseconds under load) is paid once per 8 tiles, not once per tile.
        // Groups rotate over the A decider warps.
        const int a = warp - NC - 2;
        constexpr int GD = 32 / R;                                   // tiles per group
        double *vrow = reinterpret_cast<double *>(smem + M.vrow) + a * 128;      // [row in group][code]
        const int64_t n_groups = (n_tiles + GD - 1) / GD;
        for (int64_t g = a; g < n_groups; g += A) {
            const int64_t t0 = g * GD;
            const int ng = (int)min((int64_t)GD, n_tiles - t0);
            const int64_t my_row = (tile_lo + t0) * R + lane;
            const bool have = lane < ng * R && my_row < P.n_rows;
            npc_row row;
            if (have) row = P.rows[my_row];                          // in flight while we wait
            // the grid cannot have arrived before this CTA has: sleep on the local barrier of the
            // group's last tile first (tiles are counted in order)
            {
                const int64_t tl = t0 + ng - 1;
                mbar_wait_sleep(bar_cnt + 8u * (uint32_t)(tl % Sc), (uint32_t)((tl / Sc) & 1), P.aux_sleep_ns);
            }
            double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;           // a dropped row adds +0.0: the identity
            int used = 0;
            if (have) {
                const ull *word = P.counts + my_row;
                ull v = ld_relaxed_gpu_u64(word);
                while ((v >> 56) != (ull)Gs) { __nanosleep(500); v = ld_relaxed_gpu_u64(word); }      // all slabs of this row group
                RowP rp; npc_locus rec;
                decide_row(P.pol, row, (v >> FUSED_CNT_BITS) & FUSED_CNT_MASK, v & FUSED_CNT_MASK, P.n, rp, rec);
                used = rec.used;
                if (slab_id == 0) P.log[my_row] = rec;
                if (rp.mode == MODE_DECODE) { v0 = rp.c0; v1 = rp.c1; v2 = rp.c2; v3 = rp.cm; }
                else if (rp.mode == MODE_CONST) { v0 = v1 = v2 = v3 = rp.c0; }
            }
            vrow[lane * 4 + 0] = v0; vrow[lane * 4 + 1] = v1; vrow[lane * 4 + 2] = v2; vrow[lane * 4 + 3] = v3;
            if (slab_id == 0) {
                used = __reduce_add_sync(0xffffffffu, used);
                if (lane == 0 && used) atomicAdd(P.nloci, (ull)used);
            }
            __syncwarp();
            // per tile, default order: lanes 0..15 build T01[lane] = v0[lane&3] + v1[lane>>2], lanes 16..31
            // T23 from rows 2, 3.  Exact order: no pre-summing -- lanes 0..15 copy row (lane>>2)'s four
            // contributions to bytes 32*row + 8*code of the tile's block.
            const int h = lane >> 4, e = lane & 15;
            for (int j = 0; j < ng; j++) {
                const int s = (int)((t0 + j) % Sc);
                double *tab = reinterpret_cast<double *>(smem + M.vtab) + s * 32;
                if (EXACT) {
                    if (lane < 16) tab[lane] = vrow[j * R * 4 + lane];
                } else {
                    const double *vr = vrow + (j * R + 2 * h) * 4;
                    tab[lane] = __dadd_rn(vr[e & 3], vr[4 + (e >> 2)]);
                }
            }
            __syncwarp();
            if (lane < ng) mbar_arrive(bar_lut + 8u * (uint32_t)((t0 + lane) % Sc));
            __syncwarp();
        }
    } else {
        // ================= consumers ===========================================================
        uint32_t cell[K], own[K], tailor[K];
        int valid[K];
        double acc[K][8];                      // natural sample order
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int jc = lane + 32 * (warp + NC * k);
            cell[k] = (uint32_t)jc;
            const int64_t g = c0 + jc;
            valid[k] = jc < nch ? (int)min((int64_t)8, P.n - g * 8) : 0;
            own[k] = jc < nch ? 0xFFFFFFFFu : 0u;
            tailor[k] = (jc < nch && valid[k] < 8) ? 0xF0u : 0u;
#pragma unroll
            for (int e = 0; e < 8; e++) acc[k][e] = (e < valid[k] && grp == 0) ? P.sums[g * 8 + e] : 0.0;
        }
        double *sums_out = grp == 0 ? P.sums : P.partials + (int64_t)(grp - 1) * P.n;
        const uint32_t slab = (uint32_t)P.slab_stride, islab = slab >> 1;
        const uint32_t code_hi = (sb + M.code) >> 8;
        const int nt = (int)n_tiles;
        int sr = 0, sc = 0, sa = 0;
        uint32_t ph_r = 0, ph_a = 0;
        for (int i = 0; i < nt + L; i++) {
            if (i < nt) {
                mbar_wait(bar_full + 8u * sr, ph_r);
                const uint32_t gt_mask = s_risgt[sr];
                const uint32_t d0 = sb + M.data + (uint32_t)(sr * R) * slab;
                // bits 8.. of each row's (T, r) table address; a PRMT puts the sample's entry offset below it
                int ea4[R];
                uint32_t thi4[R];
#pragma unroll
                for (int r = 0; r < R; r++) {
                    ea4[r] = s_reaidx[sr * R + r];
                    thi4[r] = code_hi + (uint32_t)min(ea4[r], (int)FUSED_CNT_TABLES - 1) * R + r;
                }
                uint32_t d02 = 0, d13 = 0, m02 = 0, m13 = 0;     // tallies: rows (0,2) / (1,3) in 16-bit halves
#pragma unroll
                for (int k = 0; k < K; k++) {
                    uint32_t B[8];                               // per sample: OR of the rows' entries; byte 1 = four 2-bit codes
                    uint32_t TA[R], TB[R];                       // per row: sum of the entries of samples 0-3 / 4-7;
                                                                 // bits 0-3 = effect alleles, bits 4-6 = missing samples
#pragma unroll
                    for (int e = 0; e < 8; e++) B[e] = 0;
#pragma unroll
                    for (int r = 0; r < R; r++) TA[r] = TB[r] = 0;
                    uint4 w[R];
                    uint32_t hi_bits = tailor[k];
#pragma unroll
                    for (int r = 0; r < R; r++) {                // all loads of the tile first: 4 independent LDS.128
                        w[r] = lds_v4(d0 + r * slab + cell[k] * 16u);
                        hi_bits |= ((w[r].x | w[r].y) | (w[r].z | w[r].w)) & 0xF8F8F8F8u;
                    }
                    if (gt_mask == (1u << R) - 1u && hi_bits == 0u) {
                        // common case, straight line: every row has genotypes and all 64 bytes are < 8:
                        // 16 folds, 32 PRMT-composed addresses, 32 independent two-byte lookups.  Entries
                        // of different rows are merged by ADDITION, two rows per IADD3: code bits are
                        // disjoint and the tally bytes (<= 16 each) cannot carry into them.
#pragma unroll
                        for (int p2 = 0; p2 < R; p2 += 2) {
                            uint32_t v[2][8];
#pragma unroll
                            for (int q = 0; q < 2; q++) {
                                const int r = p2 + q;
                                const uint32_t f[4] = { fold_offsets(w[r].x), fold_offsets(w[r].y), fold_offsets(w[r].z), fold_offsets(w[r].w) };
#pragma unroll
                                for (int e = 0; e < 4; e++) {
                                    v[q][2 * e] = lds_u16(__byte_perm(f[e], thi4[r], 0x6540));
                                    v[q][2 * e + 1] = lds_u16(__byte_perm(f[e], thi4[r], 0x6542));
                                }
                                TA[r] = (v[q][0] + v[q][1]) + (v[q][2] + v[q][3]);
                                TB[r] = (v[q][4] + v[q][5]) + (v[q][6] + v[q][7]);
                            }
#pragma unroll
                            for (int e = 0; e < 8; e++) B[e] = B[e] + v[0][e] + v[1][e];
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < R; r++) {
                            if (!((gt_mask >> r) & 1u)) continue;
                            uint32_t v[8];
                            if (((((w[r].x | w[r].y) | (w[r].z | w[r].w)) & 0xF8F8F8F8u) | tailor[k]) == 0u) {
                                const uint32_t f[4] = { fold_offsets(w[r].x), fold_offsets(w[r].y), fold_offsets(w[r].z), fold_offsets(w[r].w) };
#pragma unroll
                                for (int e = 0; e < 4; e++) {
                                    v[2 * e] = lds_u16(__byte_perm(f[e], thi4[r], 0x6540));
                                    v[2 * e + 1] = lds_u16(__byte_perm(f[e], thi4[r], 0x6542));
                                }
                            } else {                             // exact decode, same entry format
                                const int vk = own[k] ? valid[k] : 8;
                                const uint32_t ww[4] = { w[r].x, w[r].y, w[r].z, w[r].w };
#pragma unroll
                                for (int e = 0; e < 8; e++) {
                                    const uint32_t c = e < vk ? f4_slow_code((ww[e >> 1] >> ((e & 1) * 16)) & 0xFFFFu, ea4[r]) : 0u;
                                    v[e] = (c << (8 + 2 * r)) | (c == 3u ? 16u : c);
                                }
                            }
#pragma unroll
                            for (int e = 0; e < 8; e++) B[e] |= v[e];
                            TA[r] = (v[0] + v[1]) + (v[2] + v[3]);
                            TB[r] = (v[4] + v[5]) + (v[6] + v[7]);
                        }
                    }
                    // the tile's code bytes (byte 1 of each B) -> index ring
                    const uint32_t x = __byte_perm(__byte_perm(B[0], B[1], 0x0051), __byte_perm(B[2], B[3], 0x0051), 0x5410);
                    const uint32_t y = __byte_perm(__byte_perm(B[4], B[5], 0x0051), __byte_perm(B[6], B[7], 0x0051), 0x5410);
                    sts_v2(sb + M.idx + (uint32_t)sc * islab + cell[k] * 8u, x, y);
                    // tallies: byte 0 of each sum is clean (<= 72), byte 3 is zero: PRMT pairs rows into 16-bit halves
                    const uint32_t a02 = __byte_perm(TA[0], TA[2], 0x7430), b02 = __byte_perm(TB[0], TB[2], 0x7430);
                    const uint32_t a13 = __byte_perm(TA[1], TA[3], 0x7430), b13 = __byte_perm(TB[1], TB[3], 0x7430);
                    d02 += ((a02 & 0x000F000Fu) + (b02 & 0x000F000Fu)) & own[k];
                    d13 += ((a13 & 0x000F000Fu) + (b13 & 0x000F000Fu)) & own[k];
                    m02 += (((a02 >> 4) & 0x00070007u) + ((b02 >> 4) & 0x00070007u)) & own[k];
                    m13 += (((a13 >> 4) & 0x00070007u) + ((b13 >> 4) & 0x00070007u)) & own[k];
                }
                // raw stage fully read
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_rempty + 8u * sr);
                {   // sums over the warp stay below 2^16 per half
                    d02 = __reduce_add_sync(0xffffffffu, d02); d13 = __reduce_add_sync(0xffffffffu, d13);
                    m02 = __reduce_add_sync(0xffffffffu, m02); m13 = __reduce_add_sync(0xffffffffu, m13);
                    if (lane == 0) {
                        uint32_t *p = &s_cntacc[(sc * R) * 16 + warp];
                        p[0] = (d02 & 0xFFFFu) | (m02 << 16);
                        p[16] = (d13 & 0xFFFFu) | (m13 << 16);
                        p[32] = (d02 >> 16) | (m02 & 0xFFFF0000u);
                        p[48] = (d13 >> 16) | (m13 & 0xFFFF0000u);
                    }
                }
                if (warp == 0 && lane == 0) s_cisgt[sc] = gt_mask;
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_cnt + 8u * sc);
                if (++sr == Sr) { sr = 0; ph_r ^= 1u; }
                if (++sc == Sc) sc = 0;
            }
            if (i >= L) {
                mbar_wait(bar_lut + 8u * sa, ph_a);
                const uint32_t thi = (sb + M.vtab + (uint32_t)sa * 256u) >> 8;     // bits 8.. of the slot's table block
#pragma unroll
                for (int k = 0; k < K; k++) {
                    const uint2 v = lds_v2(sb + M.idx + (uint32_t)sa * islab + cell[k] * 8u);
                    const uint32_t vv[2] = { v.x, v.y };
                    if (EXACT) {
                        // the reference's chain: one rounded add per row, rows in order.  Row r's table holds
                        // its 4 contributions at bytes 32r + 8*code; a dropped row adds +0.0 (the identity)
#pragma unroll
                        for (int r = 0; r < R; r++)
#pragma unroll
                            for (int h = 0; h < 2; h++) {
                                const uint32_t sh = r == 0 ? vv[h] << 3 : r == 1 ? vv[h] << 1 : r == 2 ? vv[h] >> 1 : vv[h] >> 3;
                                const uint32_t off = (sh & 0x18181818u) | (0x20202020u * (uint32_t)r);
#pragma unroll
                                for (int e = 0; e < 4; e++)
                                    acc[k][4 * h + e] = __dadd_rn(acc[k][4 * h + e], lds_f64(__byte_perm(off, thi, 0x6540 + e)));
                            }
                    } else {
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            // byte offsets of 4 samples at once: T01 entry (b & 15) * 8, T23 entry 128 + (b >> 4) * 8
                            const uint32_t lo = (vv[h] & 0x0F0F0F0Fu) << 3;
                            const uint32_t hi = ((vv[h] >> 1) & 0x78787878u) | 0x80808080u;
#pragma unroll
                            for (int e = 0; e < 4; e++) {
                                const double x01 = lds_f64(__byte_perm(lo, thi, 0x6540 + e));
                                const double x23 = lds_f64(__byte_perm(hi, thi, 0x6540 + e));
                                acc[k][4 * h + e] = __dadd_rn(__dadd_rn(acc[k][4 * h + e], x01), x23);
                            }
                        }
                    }
                }
                if (++sa == Sc) { sa = 0; ph_a ^= 1u; }
            }
        }
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int64_t g = c0 + cell[k];
#pragma unroll
            for (int e = 0; e < 8; e++)
                if (e < valid[k]) sums_out[g * 8 + e] = acc[k][e];
        }
    }
}

// sums[s] = ((sums[s] + p0[s]) + p1[s]) + ...: the row groups' partial sums, in group (= row) order
__global__ void k_add_partials(double *__restrict__ sums, const double *__restrict__ partials, int64_t n, int n_partials) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    double a = sums[s];
    for (int g = 0; g < n_partials; g++) a = __dadd_rn(a, partials[(int64_t)g * n + s]);
    sums[s] = a;
}

}  // namespace npc
