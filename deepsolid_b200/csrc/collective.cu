// The hot path's only collective, behind the C ABI: the packed energy statistics of train.py:74-80 reduced over
// the ranks with ONE ncclAllReduce (sum) of 8 doubles on the caller's stream, and the Metropolis acceptance
// (qmc.py:360-361) the same way.  NCCL is bound at run time (dlopen of the libnccl.so.2 already mapped by the host
// process, e.g. torch's), so the library has no link-time dependency and single-GPU hosts never touch it.
#include "../../include/deepsolid_b200.h"
#include "kernels.cuh"

#include <dlfcn.h>
#include <string.h>

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;      // NCCL_UNIQUE_ID_BYTES = 128
typedef int ncclResult_t;                                   // ncclSuccess = 0
constexpr int kNcclFloat64 = 8, kNcclSum = 0;               // ncclDataType_t / ncclRedOp_t values of nccl.h

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi* nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.ok ? &api : nullptr;
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);          // the copy the host process already uses
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { ds_set_error("NCCL is not available: %s", dlerror()); return nullptr; }
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
    api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce;
    if (!api.ok) ds_set_error("libnccl.so.2 lacks one of ncclGetUniqueId / ncclCommInitRank / ncclCommDestroy / ncclAllReduce");
    return api.ok ? &api : nullptr;
}

int nccl_check(NcclApi* a, ncclResult_t r, const char* what) {
    if (r == 0) return 0;
    ds_set_error("%s failed: %s", what, a->GetErrorString ? a->GetErrorString(r) : "NCCL error");
    return DS_ERR_CUDA;
}

// raw sums of this rank -> the addends of the reduction.
//   local variance (train.py:76-80 as written: every device subtracts ITS OWN |Re mean|^2 before the pmean) or, with
//   global_variance != 0, the moments from which the variance about the global mean is formed afterwards.
__global__ void pack_stats_kernel(const double* __restrict__ s6, const double* __restrict__ n_accept, double moves,
                                  int global_variance, double* __restrict__ pk) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double n = s6[5];
    const double inv = n > 0.0 ? 1.0 / n : 0.0;
    const double mre = s6[0] * inv, mim = s6[1] * inv, m2 = s6[2] * inv;
    pk[0] = mre; pk[1] = mim;
    pk[2] = global_variance ? m2 : m2 - mre * mre;
    pk[3] = s6[3] * inv; pk[4] = s6[4] * inv;
    pk[5] = 1.0;                                            // rank count
    pk[6] = n;                                              // walkers
    pk[7] = (n_accept && moves > 0.0) ? n_accept[0] / moves : 0.0;     // this rank's pmove
}

__global__ void finish_stats_kernel(const double* __restrict__ pk, int global_variance, double* __restrict__ out8) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double r = pk[5] > 0.0 ? 1.0 / pk[5] : 0.0;
    const double loss = pk[0] * r, imag = pk[1] * r;
    out8[0] = loss; out8[1] = imag;
    out8[2] = global_variance ? pk[2] * r - loss * loss : pk[2] * r;
    out8[3] = pk[3] * r; out8[4] = pk[4] * r;
    out8[5] = pk[5]; out8[6] = pk[6];
    out8[7] = pk[7] * r;
}

}  // namespace

extern "C" int ds_nccl_unique_id(char* id128) {
    DS_REQUIRE(id128, "null argument");
    NcclApi* a = nccl();
    if (!a) return DS_ERR_UNSUPPORTED;
    ncclUniqueId id;
    if (int rc = nccl_check(a, a->GetUniqueId(&id), "ncclGetUniqueId")) return rc;
    memcpy(id128, id.internal, 128);
    return 0;
}

extern "C" int ds_nccl_comm_init(void** comm, int n_ranks, const char* id128, int rank, int device) {
    DS_REQUIRE(comm && id128 && n_ranks >= 1 && rank >= 0 && rank < n_ranks, "bad communicator arguments");
    NcclApi* a = nccl();
    if (!a) return DS_ERR_UNSUPPORTED;
    DS_CUDA_CHECK(cudaSetDevice(device));
    ncclUniqueId id;
    memcpy(id.internal, id128, 128);
    ncclComm_t c = nullptr;
    if (int rc = nccl_check(a, a->CommInitRank(&c, n_ranks, id, rank), "ncclCommInitRank")) return rc;
    *comm = (void*)c;
    return 0;
}

extern "C" int ds_nccl_comm_destroy(void* comm) {
    if (!comm) return 0;
    NcclApi* a = nccl();
    if (!a) return DS_ERR_UNSUPPORTED;
    return nccl_check(a, a->CommDestroy((ncclComm_t)comm), "ncclCommDestroy");
}

extern "C" int ds_stats_allreduce(ds_ctx* ctx, void* comm, const double* stats6_dev, const double* n_accept_dev,
                                  double moves_per_rank, int global_variance, double* out8_dev, void* stream) {
    DS_REQUIRE(ctx && stats6_dev && out8_dev, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    double* pk = ds_ctx_scratch8(ctx);
    DS_REQUIRE(pk, "no scratch");
    pack_stats_kernel<<<1, 32, 0, st>>>(stats6_dev, n_accept_dev, moves_per_rank, global_variance, pk);
    DS_CUDA_CHECK(cudaGetLastError());
    if (comm) {                                             // comm == NULL: single rank, the reduction is the identity
        NcclApi* a = nccl();
        if (!a) return DS_ERR_UNSUPPORTED;
        if (int rc = nccl_check(a, a->AllReduce(pk, pk, 8, kNcclFloat64, kNcclSum, (ncclComm_t)comm, st), "ncclAllReduce")) return rc;
    }
    finish_stats_kernel<<<1, 32, 0, st>>>(pk, global_variance, out8_dev);
    DS_CUDA_CHECK(cudaGetLastError());
    ds_ctx_count_launches(ctx, 2);
    return 0;
}
