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
//    (R x owned bytes) into a ring of S shared-memory stages with TMA bulk copies
//    (cp.async.bulk + mbarrier complete_tx).
//  * Consumer warps run two software-pipelined phases on the ring: COUNT on tile i (decode each
//    sample to a table offset, tally through an integer table, write the offsets back in place)
//    and ACCUMULATE on tile i-L (one 8-byte table load + one DADD per sample).
//  * Between the phases sits a grid-wide dependency, not a grid-wide barrier: the CTA adds its
//    per-row tallies to global counters (one 64-bit RED per row) and bumps the tile's arrival
//    counter; L tiles later an auxiliary warp waits until that counter reads gridDim.x, makes
//    the reference's fp64 decision for each row of the tile and builds the value tables.  The
//    wait is L tile-times after the arrive, so it is normally already satisfied.
//  * Rows are accumulated strictly in order, with the same rounded products as the reference:
//    sums are bit-identical to the two-kernel path and to the CPU oracle.
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
    ull *counts;               // [n_rows] packed nmiss<<32 | neff, zeroed
    unsigned *arrive;          // [n_tiles], zeroed
    npc_locus *log;
    ull *nloci;
    int32_t R, S, L;           // rows per tile, ring stages, count->accumulate lag in tiles
    int32_t nc;                // consumer warps
    int32_t slab_stride;       // bytes per row in a stage (>= max owned bytes, multiple of 16)
};

// ---- PTX helpers --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier; the slab is read
// once, so it is marked evict-first in L2.
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned *p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// shared-memory carve-up (all offsets in bytes from the dynamic smem base, 128-byte aligned)
struct FusedSmem {
    uint32_t data, lut, cnt, cntacc, mode, eaidx, bars, total;
    __host__ __device__ static FusedSmem make(int R, int S, int slab_stride) {
        FusedSmem m;
        uint32_t o = 0;
        m.bars = o;   o += 4u * S * 8u;            o = (o + 127u) & ~127u;   // full, cntdone, lutready, empty
        m.cntacc = o; o += (uint32_t)S * R * 8u;   o = (o + 127u) & ~127u;
        m.mode = o;   o += (uint32_t)S * R * 4u;   o = (o + 127u) & ~127u;
        m.eaidx = o;  o += (uint32_t)S * R * 4u;   o = (o + 127u) & ~127u;
        m.lut = o;    o += (uint32_t)S * R * LUT_N * 8u; o = (o + 127u) & ~127u;
        m.cnt = o;    o += (uint32_t)S * R * LUT_N * 8u; o = (o + 127u) & ~127u;
        m.data = o;   o += (uint32_t)S * R * (uint32_t)slab_stride;
        m.total = o;
        return m;
    }
};

// K = 16-byte chunks (8 samples each) owned per consumer thread
template <int K>
__global__ void __launch_bounds__(576, 1)      // <= 16 consumer warps + producer + decider
k_fused_i8x2(const FusedParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int R = P.R, S = P.S, L = P.L, NC = P.nc;
    const FusedSmem M = FusedSmem::make(R, S, P.slab_stride);
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(smem + M.bars);
    uint64_t *bar_cnt = bar_full + S, *bar_lut = bar_full + 2 * S, *bar_empty = bar_full + 3 * S;
    ull *s_cntacc = reinterpret_cast<ull *>(smem + M.cntacc);
    int32_t *s_mode = reinterpret_cast<int32_t *>(smem + M.mode);
    int32_t *s_eaidx = reinterpret_cast<int32_t *>(smem + M.eaidx);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_tiles = (P.n_rows + R - 1) / R;

    // balanced contiguous chunk range of this CTA
    const int64_t C = (P.n + 7) >> 3;
    const int64_t q = C / gridDim.x, rem = C % gridDim.x;
    const int64_t c0 = (int64_t)blockIdx.x * q + min((int64_t)blockIdx.x, rem);
    const int nch = (int)(q + ((int64_t)blockIdx.x < rem ? 1 : 0));
    const uint32_t slab_bytes = (uint32_t)nch * 16u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; s++) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_cnt[s], NC);
            mbar_init(&bar_lut[s], 1);
            mbar_init(&bar_empty[s], NC);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < S * R; i += blockDim.x) s_cntacc[i] = 0;
    __syncthreads();

    if (warp == NC) {
        // ================= producer: row metadata + TMA bulk loads ==========================
        const uint64_t pol = l2_evict_first_policy();
        for (int64_t t = 0; t < n_tiles; t++) {
            const int s = (int)(t % S);
            const uint32_t ph = (uint32_t)((t / S) & 1);
            mbar_wait(&bar_empty[s], ph ^ 1u);
            const int nr = (int)min((int64_t)R, P.n_rows - t * R);
            uint32_t n_gt = 0;
            for (int r = 0; r < nr; r++) {
                const npc_row row = P.rows[t * R + r];
                const bool is_gt = row.kind == NPC_KIND_GT && row.gt_row >= 0;
                if (is_gt) {
                    uint2 *cnt = reinterpret_cast<uint2 *>(smem + M.cnt) + (s * R + r) * LUT_N;
                    for (int e = lane; e < LUT_N; e += 32) {
                        const int c = lut_code(e, row.eaidx + 1);
                        cnt[e] = make_uint2(c == 3 ? 0x10000u : (uint32_t)c, 0u);
                    }
                    n_gt++;
                }
                if (lane == 0) { s_mode[s * R + r] = is_gt ? MODE_DECODE : MODE_SKIP; s_eaidx[s * R + r] = row.eaidx; }
            }
            for (int r = nr; r < R; r++) if (lane == 0) s_mode[s * R + r] = MODE_SKIP;
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_expect_tx(&bar_full[s], n_gt * slab_bytes);      // releases the table writes too
                if (slab_bytes)
                    for (int r = 0; r < nr; r++) {
                        const npc_row row = P.rows[t * R + r];
                        if (row.kind == NPC_KIND_GT && row.gt_row >= 0)
                            tma_load_1d(smem + M.data + (size_t)(s * R + r) * P.slab_stride,
                                        P.gt + (int64_t)row.gt_row * P.row_stride + c0 * 16, slab_bytes, &bar_full[s], pol);
                    }
            }
            __syncwarp();
        }
    } else if (warp == NC + 1) {
        // ================= publisher + decider ================================================
        for (int64_t i = 0; i < n_tiles + L; i++) {
            if (i < n_tiles) {                                   // publish this CTA's tallies of tile i
                const int s = (int)(i % S);
                mbar_wait(&bar_cnt[s], (uint32_t)((i / S) & 1));
                const int nr = (int)min((int64_t)R, P.n_rows - i * R);
                if (lane < nr) {
                    const ull v = s_cntacc[s * R + lane];
                    s_cntacc[s * R + lane] = 0;
                    if (v) atomicAdd(&P.counts[i * R + lane], v);
                    __threadfence();
                }
                __syncwarp();
                if (lane == 0) red_release_gpu_add(&P.arrive[i], 1u);
            }
            const int64_t j = i - L;
            if (j >= 0) {                                        // decide tile j, build its value tables
                const int s = (int)(j % S);
                if (lane == 0) while (ld_acquire_gpu(&P.arrive[j]) < gridDim.x) { }
                __syncwarp();
                const int nr = (int)min((int64_t)R, P.n_rows - j * R);
                RowP rp; rp.c0 = rp.c1 = rp.c2 = rp.cm = 0.0; rp.mode = MODE_SKIP; rp.eaidx = 0;
                int used = 0;
                if (lane < nr) {
                    const npc_row row = P.rows[j * R + lane];
                    const ull v = __ldcg(&P.counts[j * R + lane]);
                    npc_locus rec;
                    decide_row(P.pol, row, v >> 32, v & 0xFFFFFFFFull, P.n, rp, rec);
                    used = rec.used;
                    if (blockIdx.x == 0) P.log[j * R + lane] = rec;
                }
                if (blockIdx.x == 0) {
                    used = __reduce_add_sync(0xffffffffu, used);
                    if (lane == 0 && used) atomicAdd(P.nloci, (ull)used);
                }
                for (int r = 0; r < nr; r++) {
                    const double c0v = __shfl_sync(0xffffffffu, rp.c0, r), c1v = __shfl_sync(0xffffffffu, rp.c1, r);
                    const double c2v = __shfl_sync(0xffffffffu, rp.c2, r), cmv = __shfl_sync(0xffffffffu, rp.cm, r);
                    const int mode = __shfl_sync(0xffffffffu, rp.mode, r), ea = __shfl_sync(0xffffffffu, rp.eaidx, r);
                    double *lut = reinterpret_cast<double *>(smem + M.lut) + (s * R + r) * LUT_N;
                    if (mode == MODE_DECODE)
                        for (int e = lane; e < LUT_N; e += 32) {
                            const int code = lut_code(e, ea + 1);
                            lut[e] = code == 0 ? c0v : code == 1 ? c1v : code == 2 ? c2v : cmv;
                        }
                    else if (lane == 0) lut[0] = c0v;
                    if (lane == 0) s_mode[s * R + r] = mode;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_lut[s]);
            }
        }
    } else if (warp < NC) {
        // ================= consumers: count tile i, accumulate tile i-L =======================
        int jc[K]; bool own[K]; int valid[K];
        double acc[K][8];
#pragma unroll
        for (int k = 0; k < K; k++) {
            jc[k] = lane + 32 * (warp + NC * k);
            own[k] = jc[k] < nch;
            const int64_t g = c0 + jc[k];
            valid[k] = own[k] ? (int)min((int64_t)8, P.n - g * 8) : 0;
#pragma unroll
            for (int e = 0; e < 8; e++) acc[k][e] = e < valid[k] ? P.sums[g * 8 + e] : 0.0;
        }
        for (int64_t i = 0; i < n_tiles + L; i++) {
            if (i < n_tiles) {
                const int s = (int)(i % S);
                mbar_wait(&bar_full[s], (uint32_t)((i / S) & 1));
                for (int r = 0; r < R; r++) {
                    if (s_mode[s * R + r] != MODE_DECODE) continue;
                    uint8_t *slab = smem + M.data + (size_t)(s * R + r) * P.slab_stride;
                    const char *cnt = reinterpret_cast<const char *>(smem + M.cnt) + (size_t)(s * R + r) * LUT_N * 8;
                    const int ea = s_eaidx[s * R + r];
                    uint32_t tally = 0;                          // low half: effect alleles, high half: missing samples
#pragma unroll
                    for (int k = 0; k < K; k++) {
                        if (!own[k]) continue;
                        uint4 *cell = reinterpret_cast<uint4 *>(slab) + jc[k];
                        const uint4 w = *cell;
                        uint32_t ww[4] = { w.x, w.y, w.z, w.w }, o[4];
                        if (valid[k] == 8 && chunk_is_fast(w)) {
#pragma unroll
                            for (int e = 0; e < 4; e++) o[e] = pack_idx8(ww[e]);
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; e++) {
                                const uint32_t lo = 2 * e < valid[k] ? slow_off8(ww[e] & 0xFFFFu, ea) : 64u * 8u;
                                const uint32_t hi = 2 * e + 1 < valid[k] ? slow_off8(ww[e] >> 16, ea) : 64u * 8u;
                                o[e] = lo | (hi << 16);
                            }
                        }
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            tally += *reinterpret_cast<const uint32_t *>(cnt + (o[e] & 0xFFFFu));
                            tally += *reinterpret_cast<const uint32_t *>(cnt + (o[e] >> 16));
                        }
                        *cell = make_uint4(o[0], o[1], o[2], o[3]);
                    }
                    tally = __reduce_add_sync(0xffffffffu, tally);
                    if (lane == 0 && tally)
                        atomicAdd(&s_cntacc[s * R + r], ((ull)(tally >> 16) << 32) | (ull)(tally & 0xFFFFu));
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_cnt[s]);
            }
            const int64_t j = i - L;
            if (j >= 0) {
                const int s = (int)(j % S);
                mbar_wait(&bar_lut[s], (uint32_t)((j / S) & 1));
                for (int r = 0; r < R; r++) {
                    const int mode = s_mode[s * R + r];
                    const char *lut = reinterpret_cast<const char *>(smem + M.lut) + (size_t)(s * R + r) * LUT_N * 8;
                    if (mode == MODE_DECODE) {
                        const uint8_t *slab = smem + M.data + (size_t)(s * R + r) * P.slab_stride;
#pragma unroll
                        for (int k = 0; k < K; k++) {
                            if (!own[k]) continue;
                            const uint4 w = *(reinterpret_cast<const uint4 *>(slab) + jc[k]);
                            const uint32_t o[4] = { w.x, w.y, w.z, w.w };
#pragma unroll
                            for (int e = 0; e < 4; e++) {
                                acc[k][2 * e] = __dadd_rn(acc[k][2 * e], *reinterpret_cast<const double *>(lut + (o[e] & 0xFFFFu)));
                                acc[k][2 * e + 1] = __dadd_rn(acc[k][2 * e + 1], *reinterpret_cast<const double *>(lut + (o[e] >> 16)));
                            }
                        }
                    } else if (mode == MODE_CONST) {
                        const double c = *reinterpret_cast<const double *>(lut);
#pragma unroll
                        for (int k = 0; k < K; k++)
#pragma unroll
                            for (int e = 0; e < 8; e++) acc[k][e] = __dadd_rn(acc[k][e], c);
                    }
                }
                fence_proxy_async_smem();       // our in-place writes are ordered before the next TMA fill
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_empty[s]);
            }
        }
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int64_t g = c0 + jc[k];
#pragma unroll
            for (int e = 0; e < 8; e++)
                if (e < valid[k]) P.sums[g * 8 + e] = acc[k][e];
        }
    }
}

}  // namespace npc
