// Probe (development aid, not product): one CTA, D[128 x 256] (int32, TMEM) = A[128 x 128] (int8) * B[256 x 128]^T
// (int8), both K-major in the 128-byte-swizzle canonical layout, four tcgen05.mma.kind::i8 (K = 32 each).
// Checks the smem descriptors (B both K-major and MN-major = samples contiguous per K row), the instruction
// descriptor and the tcgen05.ld lane/column mapping.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_i8_probe umma_i8_probe.cu && ./umma_i8_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);          // start address
    d |= (uint64_t)0 << 16;                           // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
    return d;
}

__global__ void __launch_bounds__(128, 1) k_probe(const int8_t *A, const int8_t *B, int32_t *D, int mode, uint32_t lbo, uint32_t sbo) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sA = smem;                 // 128 rows x 128 B
    uint8_t *sB = smem + 16384;         // 256 rows x 128 B
    __shared__ uint32_t tmem_base;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // fill: row r, byte k -> r*128 + ((k>>4) ^ (r&7))*16 + (k&15)
    for (int i = tid; i < 128 * 128; i += 128) { int r = i >> 7, k = i & 127; sA[r * 128 + (((k >> 4) ^ (r & 7)) << 4) + (k & 15)] = (uint8_t)A[i]; }
    for (int i = tid; i < 256 * 128; i += 128) {
        int r = i >> 7, k = i & 127;                  // B[n = r][k]
        if (mode == 0) sB[r * 128 + (((k >> 4) ^ (r & 7)) << 4) + (k & 15)] = (uint8_t)B[i];
        else { int nb = r >> 7, nn = r & 127; sB[nb * 16384 + k * 128 + ((((nn >> 4) ^ (k & 7))) << 4) + (nn & 15)] = (uint8_t)B[i]; }   // MN-major: samples contiguous per K row
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" :: "r"(smem_u32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> async proxy (tensor core)
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tb = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(mode != 0) << 16) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
        for (int ks = 0; ks < 4; ks++) {
            uint64_t da = make_desc_sw128(smem_u32(sA) + ks * 32), db = make_desc_sw128(smem_u32(sB) + ks * 32);
            if (mode != 0) {
                const uint32_t sa = smem_u32(sB) + ks * 32 * 128;       // 32 K rows further
                db = (uint64_t)((sa >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
            }
            uint32_t acc = ks > 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                         :: "r"(tb), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
    }
    // everyone waits for the MMAs
    {
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    // warp w reads lanes 32w..32w+31, 256 columns, 32 at a time
    for (int c0 = 0; c0 < 256; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
                       "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
                       "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; j++) D[(warp * 32 + lane) * 256 + c0 + j] = (int32_t)v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" :: "r"(tb));
}

int main() {
    std::vector<int8_t> A(128 * 128), B(256 * 128);
    srand(1);
    for (auto &x : A) x = (int8_t)(rand() % 256 - 128);
    for (auto &x : B) x = (int8_t)(rand() % 4);
    int8_t *dA, *dB; int32_t *dD;
    cudaMalloc(&dA, A.size()); cudaMalloc(&dB, B.size()); cudaMalloc(&dD, 128 * 256 * 4);
    cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xFF, 128 * 256 * 4);
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152 + 1024);
    struct { int mode; uint32_t lbo, sbo; const char *what; } cfg[] = {
        {0, 0, 1024, "B K-major SW128"}, {1, 16384, 1024, "B MN-major SW128, LBO=16384 SBO=1024"} };
    int rc = 0;
    for (auto &c : cfg) {
        cudaMemset(dD, 0xFF, 128 * 256 * 4);
        k_probe<<<1, 128, 49152>>>(dA, dB, dD, c.mode, c.lbo, c.sbo);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 2; }
        std::vector<int32_t> D(128 * 256);
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        long bad = 0;
        for (int m = 0; m < 128; m++) for (int n = 0; n < 256; n++) {
            int32_t ref = 0;
            for (int k = 0; k < 128; k++) ref += (int32_t)A[m * 128 + k] * (int32_t)B[n * 128 + k];
            if (ref != D[m * 256 + n]) { if (bad < 3) printf("  mismatch m=%d n=%d ref=%d got=%d\n", m, n, ref, D[m * 256 + n]); bad++; }
        }
        printf("umma_i8_probe [%s]: %ld mismatches of %d\n", c.what, bad, 128 * 256);
        if (c.mode == 0) rc |= bad != 0;
    }
    return rc;
}
