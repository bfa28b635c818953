// npc_fused.cuh -- the roofline kernel: count -> decide -> accumulate in ONE persistent launch
// that reads every genotype byte from HBM exactly once (int8, ploidy 2).
//
// Problem: the contribution of a missing call needs the locus-wide tally (neff/ngt, and the
// --maxmis decision), so a row can only be accumulated after ALL samples of it were counted
// (src/nimpress.nim:563-583).  Two kernels would read the slab twice.
//
// Design (one CTA per SM, cooperative launch so all CTAs are co-resident):
//  * The sample axis is cut into 16-byte chunks (8 samples); CTA b owns a contiguous chunk
//    range for the whole launch and keeps those samples' fp64 sums in registers.
//  * Score rows are taken in tiles of R rows.  A producer warp streams each tile's slab
//    (R x owned bytes) into a RAW ring of Sr shared-memory stages with TMA bulk copies
//    (cp.async.bulk + mbarrier complete_tx).
//  * COUNT (consumer warps, tile i): decode every sample of the raw stage to a one-byte table
//    index, tally through a byte table, write the 8 index bytes of each chunk to the tile's slot
//    of the INDEX ring (half the size of the raw data) and release the raw stage at once.
//  * Between the phases sits a grid-wide dependency, not a grid-wide barrier.  Each row has one
//    64-bit word in global memory: [arrivals:8 | nmiss:28 | neff:28].  A publisher warp adds the
//    CTA's tallies of a row AND its arrival with a single RED on that word -- the data is the
//    flag, so no fence is needed.  Decider warps (tiles rotate over them) poll the R words of a
//    tile until the arrival byte reads gridDim.x, make the reference's fp64 decision per row and
//    build the tile's value tables.  Under a saturated memory system that chain takes several
//    microseconds; the index ring is deep enough (Sc tiles) to keep counting meanwhile.
//  * ACCUMULATE (consumer warps, tile i-L): one 8-byte table load + one DADD per sample, rows
//    strictly in order, with the same rounded products as the reference: sums are bit-identical
//    to the two-kernel path and to the CPU oracle.
#pragma once
#include "npc_kernels.cuh"

namespace npc {

struct FusedParams {
    const uint8_t *gt;
    int64_t row_stride;
    int64_t n;                 // samples of this context
    const npc_row *rows;
    int64_t n_rows;
    Policy pol;
    double *sums;
    ull *counts;               // [n_rows] arrivals<<56 | nmiss<<28 | neff, zeroed before the launch
    npc_locus *log;
    ull *nloci;
    int32_t Sr, Sc, L, A;      // raw stages, index-ring tiles, count->accumulate lag (tiles), decider warps
    int32_t nc;                // consumer warps
    int32_t slab_stride;       // bytes per row in a raw stage = nc*32*K*16 (every consumer thread has a cell)
    // tile kernel only: the grid is Gs sample slabs x Gr row groups (Gr = 1: every CTA sees every row).
    // Group g > 0 accumulates into partials[(g-1)*n ..]; k_add_partials folds them in group order.
    int32_t Gs, Gr;
    double *partials;
};

constexpr int FUSED_CNT_BITS = 28;
constexpr ull FUSED_CNT_MASK = (1ull << FUSED_CNT_BITS) - 1;
// The tally table of a row depends only on T = eaidx+1.  Fast-path codes are alleles REF..ALT6
// (T <= 7); for T >= 8 no fast-path code can match, so table 8 serves every larger T.
constexpr uint32_t FUSED_CNT_TABLES = 8;
constexpr uint32_t FUSED_CNT_STRIDE = 256;     // one-byte entries; a PRMT forms (index byte | table << 8)

// ---- PTX helpers --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity), "r"(0x989680u) : "memory");   // suspend-time hint: sleep in hardware, do not spin
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier; the slab is read
// once, so it is marked evict-first in L2.
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ ull ld_relaxed_gpu_u64(const ull *p) {
    ull v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_relaxed_gpu_add_u64(ull *p, ull v) {
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 lds_v2(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_v2(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void red_shared_add_u32(uint32_t addr, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// Two raw words (4 samples) -> one word of four 6-bit table indices, bytes = samples [0, 2, 1, 3].
// (w & 0x0E0E0E0E)*33 puts index(code0 | code1<<3) of a word's two samples at bits 6..11 and 22..27.
__device__ __forceinline__ uint32_t pack_idx_bytes(uint32_t w0, uint32_t w1) {
    const uint32_t p0 = (w0 & 0x0E0E0E0Eu) * 33u, p1 = (w1 & 0x0E0E0E0Eu) * 33u;
    return ((p0 >> 6) & 0x003F003Fu) | ((p1 << 2) & 0x3F003F00u);
}
// exact decode of one int8 diploid sample (low 16 bits of h) -> canonical index 64 + {d | 3 = missing}
__device__ __forceinline__ uint32_t slow_idx(uint32_t h, int eaidx) {
    int8_t a[2] = { (int8_t)(h & 0xFF), (int8_t)((h >> 8) & 0xFF) };
    int d; bool miss;
    decode_sample<int8_t>(a, 2, eaidx, d, miss);
    return (uint32_t)(64 + (miss ? 3 : d));
}

// shared-memory carve-up (byte offsets from the dynamic smem base)
struct FusedSmem {
    uint32_t bars, cntacc, mode, cisgt, risgt, reaidx, cnt, lut, idx, data, total;
    __host__ __device__ static FusedSmem make(int R, int Sr, int Sc, int slab_stride) {
        FusedSmem m;
        uint32_t o = 0;
        m.cnt = o;    o += FUSED_CNT_TABLES * FUSED_CNT_STRIDE;                            // first: 256-byte aligned tables
        m.bars = o;   o += (2u * Sr + 2u * Sc) * 8u;             o = (o + 127u) & ~127u;  // full, r_empty | cnt_done, lut_ready
        m.cntacc = o; o += (uint32_t)Sc * R * 16u * 4u;          o = (o + 127u) & ~127u;  // [slot][row][consumer warp]
        m.mode = o;   o += (uint32_t)Sc * R * 4u;                o = (o + 127u) & ~127u;
        m.cisgt = o;  o += (uint32_t)Sc * 4u;                    o = (o + 127u) & ~127u;  // per index slot: mask of rows with genotypes
        m.risgt = o;  o += (uint32_t)Sr * 4u;                    o = (o + 127u) & ~127u;  // per raw stage: same mask
        m.reaidx = o; o += (uint32_t)Sr * R * 4u;                o = (o + 127u) & ~127u;
        m.lut = o;    o += (uint32_t)Sc * R * LUT_N * 8u;        o = (o + 127u) & ~127u;
        m.idx = o;    o += (uint32_t)Sc * R * (uint32_t)(slab_stride / 2); o = (o + 127u) & ~127u;
        m.data = o;   o += (uint32_t)Sr * R * (uint32_t)slab_stride;
        m.total = o;
        return m;
    }
};

// K = 16-byte chunks (8 samples each) per consumer thread, R = score rows per tile
template <int K, int R>
__global__ void __launch_bounds__(768, 1)      // <= 16 consumer warps + producer + publisher + <= 6 deciders
k_fused_i8x2(const FusedParams P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int Sr = P.Sr, Sc = P.Sc, L = P.L, NC = P.nc, A = P.A;
    const FusedSmem M = FusedSmem::make(R, Sr, Sc, P.slab_stride);
    const uint32_t sb = smem_u32(smem);
    const uint32_t bar_full = sb + M.bars, bar_rempty = bar_full + 8u * Sr, bar_cnt = bar_rempty + 8u * Sr, bar_lut = bar_cnt + 8u * Sc;
    uint32_t *s_cntacc = reinterpret_cast<uint32_t *>(smem + M.cntacc);
    int32_t *s_mode = reinterpret_cast<int32_t *>(smem + M.mode);
    uint32_t *s_risgt = reinterpret_cast<uint32_t *>(smem + M.risgt);
    uint32_t *s_cisgt = reinterpret_cast<uint32_t *>(smem + M.cisgt);
    int32_t *s_reaidx = reinterpret_cast<int32_t *>(smem + M.reaidx);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_tiles = (P.n_rows + R - 1) / R;

    // balanced contiguous chunk range of this CTA
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
    for (int i = threadIdx.x; i < (int)FUSED_CNT_TABLES * LUT_N; i += blockDim.x) {
        const int c = lut_code(i % LUT_N, i / LUT_N + 1);                  // entry = d | missing << 5
        smem[M.cnt + (i / LUT_N) * FUSED_CNT_STRIDE + (i % LUT_N)] = (uint8_t)(c == 3 ? 32 : c);
    }
    // raw cells beyond this CTA's owned range are decoded too (branch-free consumers) but never
    // tallied or stored; give them defined contents once (TMA only ever writes the owned bytes)
    for (uint32_t i = threadIdx.x; i < (uint32_t)Sr * R * (uint32_t)P.slab_stride / 16u; i += blockDim.x)
        reinterpret_cast<uint4 *>(smem + M.data)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    if (warp == NC) {
        // ================= producer: row kinds + TMA bulk loads ===============================
        // Row metadata is read a group of tiles at a time, one row per lane, and the NEXT group is
        // already in flight while this one is issued (a dependent global load per tile would cap the
        // producer at one tile per memory round trip).
        const uint64_t pol = l2_evict_first_policy();
        constexpr int G = 32 / R;                                    // tiles per metadata group
        constexpr int GR = G * R;                                    // rows per group (<= 32)
        int s = 0; uint32_t ph = 0;
        npc_row nxt;
        if (lane < GR && lane < P.n_rows) nxt = P.rows[lane];
        for (int64_t t0 = 0; t0 < n_tiles; t0 += G) {
            const npc_row cur = nxt;
            const int64_t my_row = t0 * R + lane;
            const bool have = lane < GR && my_row < P.n_rows;
            if (lane < GR && my_row + GR < P.n_rows) nxt = P.rows[my_row + GR];
            const bool is_gt = have && cur.kind == NPC_KIND_GT && cur.gt_row >= 0;
            const uint32_t gt_all = __ballot_sync(0xffffffffu, is_gt);
            const int ng = (int)min((int64_t)G, n_tiles - t0);
            for (int j = 0; j < ng; j++) {
                mbar_wait(bar_rempty + 8u * s, ph ^ 1u);
                const uint32_t gt_mask = (gt_all >> (R * j)) & ((1u << R) - 1u);
                const bool mine = lane < GR && (lane / R) == j;      // lanes R*j .. R*j+R-1 own this tile's rows
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
        // ================= publisher: one RED per row = tallies + arrival ======================
        int s = 0; uint32_t ph = 0;
        for (int64_t t = 0; t < n_tiles; t++) {
            mbar_wait(bar_cnt + 8u * s, ph);
            const int nr = (int)min((int64_t)R, P.n_rows - t * R);
            if (lane < nr) {
                ull miss = 0, eff = 0;
                if ((s_cisgt[s] >> lane) & 1u)                            // row has genotypes: every consumer warp left a partial
                    for (int w = 0; w < NC; w++) {
                        const uint32_t v = s_cntacc[(s * R + lane) * 16 + w];
                        miss += v >> 16; eff += v & 0xFFFFu;
                    }
                red_relaxed_gpu_add_u64(P.counts + t * R + lane, (1ull << 56) | (miss << FUSED_CNT_BITS) | eff);
            }
            if (++s == Sc) { s = 0; ph ^= 1u; }
        }
    } else if (warp > NC + 1) {
        // ================= deciders: wait for the grid, decide, build the value tables ========
        const int a = warp - NC - 2;
        for (int64_t t = a; t < n_tiles; t += A) {
            const int s = (int)(t % Sc);
            const int nr = (int)min((int64_t)R, P.n_rows - t * R);
            RowP rp; rp.c0 = rp.c1 = rp.c2 = rp.cm = 0.0; rp.mode = MODE_SKIP; rp.eaidx = 0;
            int used = 0;
            npc_row row;
            if (lane < nr) row = P.rows[t * R + lane];               // in flight while we wait
            // the grid cannot have arrived before this CTA has: sleep on the local barrier first
            mbar_wait(bar_cnt + 8u * s, (uint32_t)((t / Sc) & 1));
            if (lane < nr) {
                const ull *word = P.counts + t * R + lane;
                ull v = ld_relaxed_gpu_u64(word);
                while ((v >> 56) != (ull)gridDim.x) { __nanosleep(500); v = ld_relaxed_gpu_u64(word); }
                npc_locus rec;
                decide_row(P.pol, row, (v >> FUSED_CNT_BITS) & FUSED_CNT_MASK, v & FUSED_CNT_MASK, P.n, rp, rec);
                used = rec.used;
                if (blockIdx.x == 0) P.log[t * R + lane] = rec;
            }
            if (blockIdx.x == 0) {
                used = __reduce_add_sync(0xffffffffu, used);
                if (lane == 0 && used) atomicAdd(P.nloci, (ull)used);
            }
#pragma unroll
            for (int r = 0; r < R; r++) {
                const double c0v = __shfl_sync(0xffffffffu, rp.c0, r), c1v = __shfl_sync(0xffffffffu, rp.c1, r);
                const double c2v = __shfl_sync(0xffffffffu, rp.c2, r), cmv = __shfl_sync(0xffffffffu, rp.cm, r);
                const int mode = __shfl_sync(0xffffffffu, rp.mode, r), ea = __shfl_sync(0xffffffffu, rp.eaidx, r);
                double *lut = reinterpret_cast<double *>(smem + M.lut) + (s * R + r) * LUT_N;
                if (mode == MODE_DECODE) {
                    for (int e = lane; e < LUT_N; e += 32) {
                        const int code = lut_code(e, ea + 1);
                        lut[e] = code == 0 ? c0v : code == 1 ? c1v : code == 2 ? c2v : cmv;
                    }
                } else if (lane == 0) lut[0] = c0v;
                if (lane == 0) s_mode[s * R + r] = mode;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_lut + 8u * s);
        }
    } else {
        // ================= consumers: count tile i, accumulate tile i-L =======================
        // thread (warp, lane) owns cell jc[k] = lane + 32*(warp + NC*k) of every row slab
        uint32_t cell[K];                      // index of the thread's cell inside a row slab
        uint32_t own[K];                       // 0xFFFFFFFF when the cell holds samples of this CTA
        uint32_t tailor[K];                    // non-zero forces the exact decode (cohort's last, partial chunk)
        int valid[K];
        double acc[K][8];                      // sums, in index-byte order: samples [0,2,1,3,4,6,5,7] of the chunk
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int jc = lane + 32 * (warp + NC * k);
            cell[k] = (uint32_t)jc;
            const int64_t g = c0 + jc;
            valid[k] = jc < nch ? (int)min((int64_t)8, P.n - g * 8) : 0;
            own[k] = jc < nch ? 0xFFFFFFFFu : 0u;
            tailor[k] = (jc < nch && valid[k] < 8) ? 0xF0u : 0u;
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const int smp = (e & 4) | ((e & 1) << 1) | ((e >> 1) & 1);          // byte position e holds sample smp
                acc[k][e] = smp < valid[k] ? P.sums[g * 8 + smp] : 0.0;
            }
        }
        const uint32_t slab = (uint32_t)P.slab_stride, islab = slab >> 1;
        const uint32_t cnt0 = sb + M.cnt;
        const int nt = (int)n_tiles;
        int sr = 0, sc = 0, sa = 0;            // ring positions: raw stage, index slot being counted / accumulated
        uint32_t ph_r = 0, ph_a = 0;
        for (int i = 0; i < nt + L; i++) {
            if (i < nt) {
                mbar_wait(bar_full + 8u * sr, ph_r);
                const uint32_t gt_mask = s_risgt[sr];
                const uint32_t d0 = sb + M.data + (uint32_t)(sr * R) * slab;
                const uint32_t x0 = sb + M.idx + (uint32_t)(sc * R) * islab;
#pragma unroll
                for (int r = 0; r < R; r++) {
                    if (!((gt_mask >> r) & 1u)) continue;
                    const int ea = s_reaidx[sr * R + r];
                    // bits 8.. of the row's tally-table address (tables are 256-byte aligned): one PRMT then
                    // forms the complete address, table bytes above the index byte -- no add
                    const uint32_t thi = (cnt0 >> 8) + (uint32_t)min(ea, (int)FUSED_CNT_TABLES - 1);
                    uint32_t tally = 0;                          // low half: effect alleles, high half: missing samples
#pragma unroll
                    for (int k = 0; k < K; k++) {
                        const uint4 w = lds_v4(d0 + r * slab + cell[k] * 16u);
                        uint32_t i0, i1;
                        if (((((w.x | w.y) | (w.z | w.w)) & 0xF0F0F0F0u) | tailor[k]) == 0u) {
                            i0 = pack_idx_bytes(w.x, w.y);
                            i1 = pack_idx_bytes(w.z, w.w);
                        } else {
                            const int vk = own[k] ? valid[k] : 8;
                            const uint32_t ww[4] = { w.x, w.y, w.z, w.w };
                            uint32_t b[8];
#pragma unroll
                            for (int e = 0; e < 8; e++)
                                b[e] = e < vk ? slow_idx((ww[e >> 1] >> ((e & 1) * 16)) & 0xFFFFu, ea) : 64u;
                            i0 = b[0] | (b[2] << 8) | (b[1] << 16) | (b[3] << 24);
                            i1 = b[4] | (b[6] << 8) | (b[5] << 16) | (b[7] << 24);
                        }
                        sts_v2(x0 + r * islab + cell[k] * 8u, i0, i1);
                        uint32_t t = 0;
#pragma unroll
                        for (int e = 0; e < 4; e++)
                            t += lds_u8(__byte_perm(i0, thi, 0x6540 + e)) + lds_u8(__byte_perm(i1, thi, 0x6540 + e));
                        tally += ((t & 31u) | ((t >> 5) << 16)) & own[k];
                    }
                    tally = __reduce_add_sync(0xffffffffu, tally);
                    if (lane == 0) s_cntacc[(sc * R + r) * 16 + warp] = tally;     // per-warp partial, summed by the publisher
                }
                if (warp == 0 && lane == 0) s_cisgt[sc] = gt_mask;
                __syncwarp();
                if (lane == 0) { mbar_arrive(bar_rempty + 8u * sr); mbar_arrive(bar_cnt + 8u * sc); }
                if (++sr == Sr) { sr = 0; ph_r ^= 1u; }
                if (++sc == Sc) sc = 0;
            }
            if (i >= L) {
                mbar_wait(bar_lut + 8u * sa, ph_a);
                const uint32_t x0 = sb + M.idx + (uint32_t)(sa * R) * islab;
                const uint32_t l0 = sb + M.lut + (uint32_t)(sa * R) * (LUT_N * 8u);
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const int mode = s_mode[sa * R + r];
                    const uint32_t lut = l0 + r * (LUT_N * 8u);
                    if (mode == MODE_DECODE) {
#pragma unroll
                        for (int k = 0; k < K; k++) {
                            const uint2 v = lds_v2(x0 + r * islab + cell[k] * 8u);
#pragma unroll
                            for (int e = 0; e < 4; e++) {
                                acc[k][e] = __dadd_rn(acc[k][e], lds_f64(lut + (__byte_perm(v.x, 0, 0x4440 + e) << 3)));
                                acc[k][4 + e] = __dadd_rn(acc[k][4 + e], lds_f64(lut + (__byte_perm(v.y, 0, 0x4440 + e) << 3)));
                            }
                        }
                    } else if (mode == MODE_CONST) {
                        const double c = lds_f64(lut);
#pragma unroll
                        for (int k = 0; k < K; k++)
#pragma unroll
                            for (int e = 0; e < 8; e++) acc[k][e] = __dadd_rn(acc[k][e], c);
                    }
                }
                if (++sa == Sc) { sa = 0; ph_a ^= 1u; }
            }
        }
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int64_t g = c0 + cell[k];
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const int smp = (e & 4) | ((e & 1) << 1) | ((e >> 1) & 1);
                if (smp < valid[k]) P.sums[g * 8 + smp] = acc[k][e];
            }
        }
    }
}

}  // namespace npc
