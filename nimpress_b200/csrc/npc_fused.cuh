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
//  * Between the phases sits a grid-wide dependency, not a grid-wide barrier.  Each row has one
//    64-bit word in global memory: [arrivals:8 | nmiss:28 | neff:28].  A CTA publishes its
//    tallies of a row AND its arrival with a single RED on that word -- the data is the flag,
//    so no fence is needed.  An auxiliary warp that owns the tile then polls the R words until
//    the arrival byte reads gridDim.x, makes the reference's fp64 decision per row and builds the
//    value tables.  Tiles rotate over A auxiliary warps so their latency chains overlap, and the
//    accumulate phase runs L tiles behind the count phase so the chain is off the critical path.
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
    ull *counts;               // [n_rows] arrivals<<56 | nmiss<<28 | neff, zeroed before the launch
    npc_locus *log;
    ull *nloci;
    int32_t S, L, A;           // ring stages, count->accumulate lag in tiles, auxiliary warps
    int32_t nc;                // consumer warps
    int32_t slab_stride;       // bytes per row in a stage = nc*32*K*16 (every consumer thread has a cell)
};

constexpr int FUSED_CNT_BITS = 28;
// The tally table of a row depends only on T = eaidx+1.  Fast-path codes are alleles REF..ALT6
// (T <= 7); for T >= 8 no fast-path code can match, so table 8 serves every larger T.
constexpr uint32_t FUSED_CNT_TABLES = 8;
constexpr ull FUSED_CNT_MASK = (1ull << FUSED_CNT_BITS) - 1;

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
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
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
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ ull ld_relaxed_gpu_u64(const ull *p) {
    ull v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_relaxed_gpu_add_u64(ull *p, ull v) {
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
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
__device__ __forceinline__ void sts_v4(uint32_t addr, const uint4 &v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// shared-memory carve-up (byte offsets from the dynamic smem base)
struct FusedSmem {
    uint32_t bars, cntacc, mode, eaidx, lut, cnt, data, total;
    __host__ __device__ static FusedSmem make(int R, int S, int slab_stride) {
        FusedSmem m;
        uint32_t o = 0;
        m.bars = o;   o += 4u * S * 8u;                  o = (o + 127u) & ~127u;   // full, cntdone, lutready, empty
        m.cntacc = o; o += (uint32_t)S * R * 4u;         o = (o + 127u) & ~127u;
        m.mode = o;   o += (uint32_t)S * R * 4u;         o = (o + 127u) & ~127u;
        m.eaidx = o;  o += (uint32_t)S * R * 4u;         o = (o + 127u) & ~127u;
        m.lut = o;    o += (uint32_t)S * R * LUT_N * 8u; o = (o + 127u) & ~127u;
        m.cnt = o;    o += FUSED_CNT_TABLES * LUT_N * 8u; o = (o + 127u) & ~127u;   // tally tables, one per T = eaidx+1
        m.data = o;   o += (uint32_t)S * R * (uint32_t)slab_stride;
        m.total = o;
        return m;
    }
};

// K = 16-byte chunks (8 samples each) per consumer thread, R = score rows per tile
template <int K, int R>
__global__ void __launch_bounds__(704, 1)      // <= 16 consumer warps + producer + <= 5 auxiliary warps
k_fused_i8x2(const FusedParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int S = P.S, L = P.L, NC = P.nc, A = P.A;
    const FusedSmem M = FusedSmem::make(R, S, P.slab_stride);
    const uint32_t sb = smem_u32(smem);
    const uint32_t bar_full = sb + M.bars, bar_cnt = bar_full + 8u * S, bar_lut = bar_full + 16u * S, bar_empty = bar_full + 24u * S;
    uint32_t *s_cntacc = reinterpret_cast<uint32_t *>(smem + M.cntacc);
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
            mbar_init(bar_full + 8u * s, 1);
            mbar_init(bar_cnt + 8u * s, NC);
            mbar_init(bar_lut + 8u * s, 1);
            mbar_init(bar_empty + 8u * s, NC);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < S * R; i += blockDim.x) s_cntacc[i] = 0;
    for (int i = threadIdx.x; i < (int)FUSED_CNT_TABLES * LUT_N; i += blockDim.x) {
        const int c = lut_code(i % LUT_N, i / LUT_N + 1);
        reinterpret_cast<uint2 *>(smem + M.cnt)[i] = make_uint2(c == 3 ? 0x10000u : (uint32_t)c, 0u);
    }
    // cells beyond this CTA's owned range are decoded too (branch-free consumers) but never tallied
    // or stored; give them defined contents once
    for (uint32_t i = threadIdx.x; i < (uint32_t)S * R * (uint32_t)P.slab_stride / 16u; i += blockDim.x)
        reinterpret_cast<uint4 *>(smem + M.data)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    __syncthreads();

    if (warp == NC) {
        // ================= producer: row modes + TMA bulk loads ===============================
        const uint64_t pol = l2_evict_first_policy();
        for (int64_t t = 0; t < n_tiles; t++) {
            const int s = (int)(t % S);
            mbar_wait(bar_empty + 8u * s, (uint32_t)((t / S) & 1) ^ 1u);
            const int nr = (int)min((int64_t)R, P.n_rows - t * R);
            uint32_t n_gt = 0;
#pragma unroll
            for (int r = 0; r < R; r++) {
                bool is_gt = false;
                int ea = 0;
                if (r < nr) {
                    const npc_row row = P.rows[t * R + r];
                    is_gt = row.kind == NPC_KIND_GT && row.gt_row >= 0;
                    ea = row.eaidx;
                    if (is_gt) n_gt++;
                }
                if (lane == 0) { s_mode[s * R + r] = is_gt ? MODE_DECODE : MODE_SKIP; s_eaidx[s * R + r] = ea; }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_expect_tx(bar_full + 8u * s, n_gt * slab_bytes);     // releases the mode writes too
                for (int r = 0; r < nr; r++) {
                    const npc_row row = P.rows[t * R + r];
                    if (row.kind == NPC_KIND_GT && row.gt_row >= 0)
                        tma_load_1d(sb + M.data + (uint32_t)(s * R + r) * (uint32_t)P.slab_stride,
                                    P.gt + (int64_t)row.gt_row * P.row_stride + c0 * 16, slab_bytes, bar_full + 8u * s, pol);
                }
            }
            __syncwarp();
        }
    } else if (warp > NC) {
        // ================= auxiliary warps: publish tallies, wait for the grid, decide, build tables
        const int a = warp - NC - 1;
        for (int64_t t = a; t < n_tiles; t += A) {
            const int s = (int)(t % S);
            const uint32_t ph = (uint32_t)((t / S) & 1);
            const int nr = (int)min((int64_t)R, P.n_rows - t * R);
            mbar_wait(bar_cnt + 8u * s, ph);
            ull *word = P.counts + t * R + lane;
            if (lane < nr) {
                const uint32_t v = s_cntacc[s * R + lane];
                s_cntacc[s * R + lane] = 0;
                red_relaxed_gpu_add_u64(word, (1ull << 56) | ((ull)(v >> 16) << FUSED_CNT_BITS) | (ull)(v & 0xFFFFu));
            }
            RowP rp; rp.c0 = rp.c1 = rp.c2 = rp.cm = 0.0; rp.mode = MODE_SKIP; rp.eaidx = 0;
            int used = 0;
            if (lane < nr) {
                ull v = ld_relaxed_gpu_u64(word);
                while ((v >> 56) != (ull)gridDim.x) { __nanosleep(40); v = ld_relaxed_gpu_u64(word); }
                const npc_row row = P.rows[t * R + lane];
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
        // thread (warp, lane) owns cells jc[k] = lane + 32*(warp + NC*k) of every row slab
        uint32_t cell[K];                      // byte offset of the thread's cell inside a row slab
        uint32_t own[K];                       // 0xFFFFFFFF when the cell holds samples of this CTA
        uint32_t tailor[K];                    // non-zero forces the exact decode (cohort's last, partial chunk)
        int valid[K];
        double acc[K][8];
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int jc = lane + 32 * (warp + NC * k);
            cell[k] = (uint32_t)jc * 16u;
            const int64_t g = c0 + jc;
            valid[k] = jc < nch ? (int)min((int64_t)8, P.n - g * 8) : 0;
            own[k] = jc < nch ? 0xFFFFFFFFu : 0u;
            tailor[k] = (jc < nch && valid[k] < 8) ? 0xF0u : 0u;
#pragma unroll
            for (int e = 0; e < 8; e++) acc[k][e] = e < valid[k] ? P.sums[g * 8 + e] : 0.0;
        }
        const uint32_t slab = (uint32_t)P.slab_stride;
        int s_c = 0, s_a = 0;                  // ring positions of the count / accumulate phase
        uint32_t ph_c = 0, ph_a = 0;
        for (int64_t i = 0; i < n_tiles + L; i++) {
            if (i < n_tiles) {
                mbar_wait(bar_full + 8u * s_c, ph_c);
                const uint32_t d0 = sb + M.data + (uint32_t)(s_c * R) * slab;
#pragma unroll
                for (int r = 0; r < R; r++) {
                    if (s_mode[s_c * R + r] != MODE_DECODE) continue;
                    const int ea = s_eaidx[s_c * R + r];
                    const uint32_t cnt = sb + M.cnt + (uint32_t)min(ea, (int)FUSED_CNT_TABLES - 1) * (LUT_N * 8u);
                    uint32_t tally = 0;                          // low half: effect alleles, high half: missing samples
#pragma unroll
                    for (int k = 0; k < K; k++) {
                        const uint32_t addr = d0 + r * slab + cell[k];
                        const uint4 w = lds_v4(addr);
                        uint32_t ww[4] = { w.x, w.y, w.z, w.w }, o[4];
                        if (((((w.x | w.y) | (w.z | w.w)) & 0xF0F0F0F0u) | tailor[k]) == 0u) {
#pragma unroll
                            for (int e = 0; e < 4; e++) o[e] = pack_idx8(ww[e]);
                        } else {
                            const int vk = own[k] ? valid[k] : 8;
#pragma unroll
                            for (int e = 0; e < 4; e++) {
                                const uint32_t lo = 2 * e < vk ? slow_off8(ww[e] & 0xFFFFu, ea) : 64u * 8u;
                                const uint32_t hi = 2 * e + 1 < vk ? slow_off8(ww[e] >> 16, ea) : 64u * 8u;
                                o[e] = lo | (hi << 16);
                            }
                        }
                        uint32_t t = 0;
#pragma unroll
                        for (int e = 0; e < 4; e++) t += lds_u32(cnt + (o[e] & 0xFFFFu)) + lds_u32(cnt + (o[e] >> 16));
                        tally += t & own[k];
                        sts_v4(addr, make_uint4(o[0], o[1], o[2], o[3]));
                    }
                    tally = __reduce_add_sync(0xffffffffu, tally);
                    if (lane == 0 && tally) atomicAdd(&s_cntacc[s_c * R + r], tally);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_cnt + 8u * s_c);
                if (++s_c == S) { s_c = 0; ph_c ^= 1u; }
            }
            if (i >= L) {
                mbar_wait(bar_lut + 8u * s_a, ph_a);
                const uint32_t d0 = sb + M.data + (uint32_t)(s_a * R) * slab;
                const uint32_t l0 = sb + M.lut + (uint32_t)(s_a * R) * (LUT_N * 8u);
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const int mode = s_mode[s_a * R + r];
                    const uint32_t lut = l0 + r * (LUT_N * 8u);
                    if (mode == MODE_DECODE) {
#pragma unroll
                        for (int k = 0; k < K; k++) {
                            const uint4 w = lds_v4(d0 + r * slab + cell[k]);
                            const uint32_t o[4] = { w.x, w.y, w.z, w.w };
#pragma unroll
                            for (int e = 0; e < 4; e++) {
                                acc[k][2 * e] = __dadd_rn(acc[k][2 * e], lds_f64(lut + (o[e] & 0xFFFFu)));
                                acc[k][2 * e + 1] = __dadd_rn(acc[k][2 * e + 1], lds_f64(lut + (o[e] >> 16)));
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
                fence_proxy_async_smem();       // our in-place writes are ordered before the next TMA fill
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_empty + 8u * s_a);
                if (++s_a == S) { s_a = 0; ph_a ^= 1u; }
            }
        }
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int64_t g = c0 + (lane + 32 * (warp + NC * k));
#pragma unroll
            for (int e = 0; e < 8; e++)
                if (e < valid[k]) P.sums[g * 8 + e] = acc[k][e];
        }
    }
}

}  // namespace npc
