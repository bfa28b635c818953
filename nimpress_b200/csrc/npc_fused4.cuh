// npc_fused4.cuh -- the default roofline kernel: the fused persistent design of npc_fused.cuh
// (TMA raw ring -> count -> grid-wide tally dependency -> decide -> accumulate, every genotype
// byte read from HBM once) restructured around TILES OF FOUR SCORE ROWS so that the per-genotype
// instruction and shared-memory cost drops by about 40%:
//
//  * COUNT.  A sample's two raw GT bytes (both < 16 on the fast path) are folded into one byte
//    b0 | b1<<4 with a single IMAD.HI per word, and a PRMT composes that byte with the
//    (effect allele, row-in-tile) table number into a complete shared-memory address: one
//    LDS.U8 returns the sample's 2-bit dosage code {0,1,2,3=missing} already shifted to the row's
//    position.  OR-ing the four rows gives ONE byte per sample per tile.
//  * The tile's tallies come from that byte through a row-independent 256-entry table
//    (per row: dosage | missing<<4, summed over 4 samples per register), once per tile.
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
// 1e-9).  The result is deterministic and independent of the launch shape.  npc_set_exact_order
// selects the bit-for-bit kernel of npc_fused.cuh instead.
#pragma once
#include "npc_fused.cuh"

namespace npc {

constexpr int F4_R = 4;                                  // rows per tile
constexpr uint32_t F4_CODE_TABLES = FUSED_CNT_TABLES * F4_R;   // (T, row-in-tile) -> 256 one-byte entries

struct Fused4Smem {
    uint32_t code, tt, vrow, bars, cntacc, cisgt, risgt, reaidx, vtab, idx, data, total;
    __host__ __device__ static Fused4Smem make(int Sr, int Sc, int slab_stride) {
        Fused4Smem m;
        uint32_t o = 0;
        m.code = o;   o += F4_CODE_TABLES * 256u;                        // first: 256-byte aligned
        m.tt = o;     o += 256u * 4u;
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
// w + (w >> 4) in one FMA-pipe op: byte 0 = b0 | b1<<4, byte 2 = b2 | b3<<4 when every byte < 16
__device__ __forceinline__ uint32_t fold_nibbles(uint32_t w) { return __umulhi(w, 0x10000000u) + w; }

template <int K>
__global__ void __launch_bounds__(768, 1)
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
    const int64_t n_tiles = (P.n_rows + R - 1) / R;

    const int64_t C = (P.n + 7) >> 3;
    const int64_t q = C / gridDim.x, rem = C % gridDim.x;
    const int64_t c0 = (int64_t)blockIdx.x * q + min((int64_t)blockIdx.x, rem);
    const int nch = (int)(q + ((int64_t)blockIdx.x < rem ? 1 : 0));
    const uint32_t slab_bytes = (uint32_t)nch * 16u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < Sr; s++) { mbar_init(bar_full + 8u * s, 1); mbar_init(bar_rempty + 8u * s, NC); }
        for (int s = 0; s < Sc; s++) { mbar_init(bar_cnt + 8u * s, NC); mbar_init(bar_lut + 8u * s, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // code tables: [T-1][row r][byte b0 | b1<<4] = dosage code << 2r
    for (uint32_t i = threadIdx.x; i < F4_CODE_TABLES * 256u; i += blockDim.x) {
        const int b = i & 255, r = (i >> 8) & 3, T = (int)(i >> 10) + 1;
        smem[M.code + i] = (uint8_t)(f4_code_from_nibbles(b & 15, b >> 4, T) << (2 * r));
    }
    // tally table: byte of four codes -> per row (dosage | missing << 4) in byte r
    for (uint32_t b = threadIdx.x; b < 256u; b += blockDim.x) {
        uint32_t v = 0;
        for (int r = 0; r < 4; r++) {
            const uint32_t c = (b >> (2 * r)) & 3u;
            v |= (c == 3u ? 16u : c) << (8 * r);
        }
        reinterpret_cast<uint32_t *>(smem + M.tt)[b] = v;
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
        npc_row nxt;
        {
            const int64_t r = lane;
            if (r < P.n_rows) nxt = P.rows[r];
        }
        for (int64_t t0 = 0; t0 < n_tiles; t0 += G) {
            const npc_row cur = nxt;
            const int64_t my_row = t0 * R + lane;
            const bool have = my_row < P.n_rows;
            {
                const int64_t r = my_row + 32;
                if (r < P.n_rows) nxt = P.rows[r];
            }
            const bool is_gt = have && cur.kind == NPC_KIND_GT && cur.gt_row >= 0;
            const uint32_t gt_all = __ballot_sync(0xffffffffu, is_gt);
            const int ng = (int)min((int64_t)G, n_tiles - t0);
            for (int j = 0; j < ng; j++) {
                mbar_wait(bar_rempty + 8u * s, ph ^ 1u);
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
            mbar_wait(bar_cnt + 8u * s, ph);
            const int nr = (int)min((int64_t)R, P.n_rows - t * R);
            if (lane < nr) {
                ull miss = 0, eff = 0;
                if ((s_cisgt[s] >> lane) & 1u)
                    for (int w = 0; w < NC; w++) {
                        const uint32_t v = s_cntacc[(s * R + lane) * 16 + w];
                        miss += v >> 16; eff += v & 0xFFFFu;
                    }
                red_relaxed_gpu_add_u64(P.counts + t * R + lane, (1ull << 56) | (miss << FUSED_CNT_BITS) | eff);
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
            const int64_t my_row = t0 * R + lane;
            const bool have = my_row < P.n_rows;
            npc_row row;
            if (have) row = P.rows[my_row];                          // in flight while we wait
            // the grid cannot have arrived before this CTA has: sleep on the local barrier of the
            // group's last tile first (tiles are counted in order)
            {
                const int64_t tl = t0 + ng - 1;
                mbar_wait(bar_cnt + 8u * (uint32_t)(tl % Sc), (uint32_t)((tl / Sc) & 1));
            }
            double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;           // a dropped row adds +0.0: the identity
            int used = 0;
            if (have) {
                const ull *word = P.counts + my_row;
                ull v = ld_relaxed_gpu_u64(word);
                while ((v >> 56) != (ull)gridDim.x) { __nanosleep(500); v = ld_relaxed_gpu_u64(word); }
                RowP rp; npc_locus rec;
                decide_row(P.pol, row, (v >> FUSED_CNT_BITS) & FUSED_CNT_MASK, v & FUSED_CNT_MASK, P.n, rp, rec);
                used = rec.used;
                if (blockIdx.x == 0) P.log[my_row] = rec;
                if (rp.mode == MODE_DECODE) { v0 = rp.c0; v1 = rp.c1; v2 = rp.c2; v3 = rp.cm; }
                else if (rp.mode == MODE_CONST) { v0 = v1 = v2 = v3 = rp.c0; }
            }
            vrow[lane * 4 + 0] = v0; vrow[lane * 4 + 1] = v1; vrow[lane * 4 + 2] = v2; vrow[lane * 4 + 3] = v3;
            if (blockIdx.x == 0) {
                used = __reduce_add_sync(0xffffffffu, used);
                if (lane == 0 && used) atomicAdd(P.nloci, (ull)used);
            }
            __syncwarp();
            // per tile: lanes 0..15 build T01[lane] = v0[lane&3] + v1[lane>>2], lanes 16..31 T23 from rows 2, 3
            const int h = lane >> 4, e = lane & 15;
            for (int j = 0; j < ng; j++) {
                const int s = (int)((t0 + j) % Sc);
                double *tab = reinterpret_cast<double *>(smem + M.vtab) + s * 32;
                const double *vr = vrow + (j * R + 2 * h) * 4;
                tab[lane] = __dadd_rn(vr[e & 3], vr[4 + (e >> 2)]);
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
            for (int e = 0; e < 8; e++) acc[k][e] = e < valid[k] ? P.sums[g * 8 + e] : 0.0;
        }
        const uint32_t slab = (uint32_t)P.slab_stride, islab = slab >> 1;
        const uint32_t code_hi = (sb + M.code) >> 8, tt0 = sb + M.tt;
        const int nt = (int)n_tiles;
        int sr = 0, sc = 0, sa = 0;
        uint32_t ph_r = 0, ph_a = 0;
        for (int i = 0; i < nt + L; i++) {
            if (i < nt) {
                mbar_wait(bar_full + 8u * sr, ph_r);
                const uint32_t gt_mask = s_risgt[sr];
                const uint32_t d0 = sb + M.data + (uint32_t)(sr * R) * slab;
                uint32_t B[K][8];                                // per sample: the four rows' codes, 2 bits each
#pragma unroll
                for (int k = 0; k < K; k++)
#pragma unroll
                    for (int e = 0; e < 8; e++) B[k][e] = 0;
                // bits 8.. of each row's (T, r) code table address; a PRMT puts the folded sample byte below it
                int ea4[R];
                uint32_t thi4[R];
#pragma unroll
                for (int r = 0; r < R; r++) {
                    ea4[r] = s_reaidx[sr * R + r];
                    thi4[r] = code_hi + (uint32_t)min(ea4[r], (int)FUSED_CNT_TABLES - 1) * R + r;
                }
#pragma unroll
                for (int k = 0; k < K; k++) {
                    uint4 w[R];
                    uint32_t hi_bits = tailor[k];
#pragma unroll
                    for (int r = 0; r < R; r++) {                // all loads of the tile first: 4 independent LDS.128
                        w[r] = lds_v4(d0 + r * slab + cell[k] * 16u);
                        hi_bits |= ((w[r].x | w[r].y) | (w[r].z | w[r].w)) & 0xF0F0F0F0u;
                    }
                    if (gt_mask == (1u << R) - 1u && hi_bits == 0u) {
                        // common case, straight line: every row has genotypes and all 64 bytes are < 16:
                        // 16 folds, 32 PRMT-composed addresses, 32 independent one-byte lookups
#pragma unroll
                        for (int r = 0; r < R; r++) {
                            const uint32_t f[4] = { fold_nibbles(w[r].x), fold_nibbles(w[r].y), fold_nibbles(w[r].z), fold_nibbles(w[r].w) };
#pragma unroll
                            for (int e = 0; e < 4; e++) {
                                B[k][2 * e] |= lds_u8(__byte_perm(f[e], thi4[r], 0x6540));
                                B[k][2 * e + 1] |= lds_u8(__byte_perm(f[e], thi4[r], 0x6542));
                            }
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < R; r++) {
                            if (!((gt_mask >> r) & 1u)) continue;
                            if (((((w[r].x | w[r].y) | (w[r].z | w[r].w)) & 0xF0F0F0F0u) | tailor[k]) == 0u) {
                                const uint32_t f[4] = { fold_nibbles(w[r].x), fold_nibbles(w[r].y), fold_nibbles(w[r].z), fold_nibbles(w[r].w) };
#pragma unroll
                                for (int e = 0; e < 4; e++) {
                                    B[k][2 * e] |= lds_u8(__byte_perm(f[e], thi4[r], 0x6540));
                                    B[k][2 * e + 1] |= lds_u8(__byte_perm(f[e], thi4[r], 0x6542));
                                }
                            } else {
                                const int vk = own[k] ? valid[k] : 8;
                                const uint32_t ww[4] = { w[r].x, w[r].y, w[r].z, w[r].w };
#pragma unroll
                                for (int e = 0; e < 8; e++)
                                    if (e < vk) B[k][e] |= f4_slow_code((ww[e >> 1] >> ((e & 1) * 16)) & 0xFFFFu, ea4[r]) << (2 * r);
                            }
                        }
                    }
                }
                // raw stage fully read
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_rempty + 8u * sr);
                // tile tallies from the code bytes (row-independent table), then one store of the bytes
                uint32_t dd = 0, mm = 0;                         // per row in byte r: effect alleles / missing samples
#pragma unroll
                for (int k = 0; k < K; k++) {
                    uint32_t ta = 0, tb = 0;
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        ta += lds_u32(tt0 + (B[k][e] << 2));
                        tb += lds_u32(tt0 + (B[k][4 + e] << 2));
                    }
                    dd += ((ta & 0x0F0F0F0Fu) + (tb & 0x0F0F0F0Fu)) & own[k];
                    mm += (((ta >> 4) & 0x07070707u) + ((tb >> 4) & 0x07070707u)) & own[k];
                    const uint32_t x = B[k][0] | (B[k][1] << 8) | (B[k][2] << 16) | (B[k][3] << 24);
                    const uint32_t y = B[k][4] | (B[k][5] << 8) | (B[k][6] << 16) | (B[k][7] << 24);
                    sts_v2(sb + M.idx + (uint32_t)sc * islab + cell[k] * 8u, x, y);
                }
                {   // rows (0,2) and (1,3) share a register as 16-bit halves: sums over the warp stay below 2^16
                    const uint32_t d02 = __reduce_add_sync(0xffffffffu, dd & 0x00FF00FFu), d13 = __reduce_add_sync(0xffffffffu, (dd >> 8) & 0x00FF00FFu);
                    const uint32_t m02 = __reduce_add_sync(0xffffffffu, mm & 0x00FF00FFu), m13 = __reduce_add_sync(0xffffffffu, (mm >> 8) & 0x00FF00FFu);
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
                if (++sa == Sc) { sa = 0; ph_a ^= 1u; }
            }
        }
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int64_t g = c0 + cell[k];
#pragma unroll
            for (int e = 0; e < 8; e++)
                if (e < valid[k]) P.sums[g * 8 + e] = acc[k][e];
        }
    }
}

}  // namespace npc
