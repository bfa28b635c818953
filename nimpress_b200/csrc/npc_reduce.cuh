// npc_reduce.cuh -- the cross-GPU combine of a variant-sharded run (included by npc_api.cu).
//
// The reference adds every locus to every sample's sum in score-file order in one process
// (src/nimpress.nim:634-641) and normalises once (:643-649).  With the score rows partitioned into
// contiguous ranges over several GPUs, each context holds the raw partial sums of its range; the
// combine is  total[s] = ((p_0[s] + p_1[s]) + p_2[s]) + ...  in context (= score-file) order -- a fixed
// order, so the result does not depend on the transport -- an integer sum of nloci, and one normalise.
//
//  * npc_reduce: one process, one context per GPU.  ONE kernel on the first context's device reads
//    the other devices' partial sums straight through NVLink peer mappings (no staging copy, no
//    second pass) and writes the normalised scores.  Devices without peer access are bridged with
//    cudaMemcpyPeerAsync into a scratch buffer first.
//  * npc_comm_*: one process per GPU (torchrun, MPI ...).  NCCL is loaded at run time (dlopen of
//    libnccl.so.2 -- the copy already in the process if the host framework brought one): all-gather of
//    the partial sums, the same fixed-order add on every rank, all-reduce of nloci.
#pragma once
#include <dlfcn.h>

namespace npc {

constexpr int REDUCE_MAX = 16;
struct ReduceParams {
    const double *part[REDUCE_MAX];      // partial sums, one per shard, in shard order
    const ull *nloci[REDUCE_MAX];        // each shard's nloci (may be peer pointers); nullptr: use nloci_total
    int n_parts;
    int normalise;                       // 1: out = total / (2 * nloci) + offset, 0: raw total
    double offset;
    ull nloci_total;
};

__global__ void __launch_bounds__(256)
k_combine(const ReduceParams P, int64_t n, double *__restrict__ out, ull *__restrict__ nloci_out) {
    ull nl = P.nloci_total;
    if (P.nloci[0]) { nl = 0; for (int k = 0; k < P.n_parts; k++) nl += *P.nloci[k]; }
    if (blockIdx.x == 0 && threadIdx.x == 0 && nloci_out) *nloci_out = nl;
    const double denom = __dmul_rn((double)(int64_t)nl, 2.0);
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (int64_t)gridDim.x * blockDim.x) {
        double a = P.part[0][s];
        for (int k = 1; k < P.n_parts; k++) a = __dadd_rn(a, P.part[k][s]);       // shard order = score-file order
        out[s] = P.normalise ? __dadd_rn(__ddiv_rn(a, denom), P.offset) : a;
    }
}

// the same with nloci already summed on the device (*nloci_total, e.g. by an all-reduce)
__global__ void __launch_bounds__(256)
k_combine_total(const ReduceParams P, int64_t n, double *__restrict__ out, const ull *__restrict__ nloci_total) {
    const double denom = __dmul_rn((double)(int64_t)*nloci_total, 2.0);
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (int64_t)gridDim.x * blockDim.x) {
        double a = P.part[0][s];
        for (int k = 1; k < P.n_parts; k++) a = __dadd_rn(a, P.part[k][s]);
        out[s] = P.normalise ? __dadd_rn(__ddiv_rn(a, denom), P.offset) : a;
    }
}

// ---- NCCL, bound at run time ------------------------------------------------------------------------
struct NcclId { char internal[128]; };
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(NcclId *) = nullptr;
    int (*CommInitRank)(void **, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string err;
    bool load() {
        if (lib) return true;
        const char *names[] = { "libnccl.so.2", "libnccl.so" };
        for (const char *nm : names) if ((lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL))) break;
        if (!lib) { err = std::string("NCCL not found (dlopen libnccl.so.2): ") + dlerror(); return false; }
        GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather");
        AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllGather || !AllReduce || !GetErrorString) {
            err = "libnccl.so.2 lacks an expected symbol"; lib = nullptr; return false;
        }
        return true;
    }
};
constexpr int NCCL_UINT64 = 5, NCCL_FLOAT64 = 8, NCCL_SUM = 0;   // ncclDataType_t / ncclRedOp_t values (nccl.h)

}  // namespace npc
