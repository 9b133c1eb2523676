// dpr_comm.cu - the multi-GPU entry points of the C ABI (include/dpr.h): one process per GPU, pose-sharded batch,
// ONE all-reduce (sum) of the packed [d_points; d_point_weight] buffer over NCCL / NVLink.
// NCCL is resolved at run time (dlopen of libnccl.so.2 - the copy the host process already loaded, e.g. the one
// bundled with PyTorch or NCCL.jl's artifact), so libdpr.so has no link-time dependency on it.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstring>
#include <mutex>

#include "dpr_internal.h"

namespace dpr {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

static NcclApi& nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) return;
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.handle, "ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.handle, "ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.handle, "ncclCommDestroy"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(api.handle, "ncclAllReduce"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.handle, "ncclGetErrorString"));
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.GetErrorString;
    });
    return api;
}

static int nccl_fail(ncclResult_t r, const char* what) {
    NcclApi& api = nccl_api();
    char buf[400];
    snprintf(buf, sizeof(buf), "%s: %s", what, api.ok ? api.GetErrorString(r) : "NCCL not available");
    set_error_message(buf);
    return DPR_ERR_NCCL;
}

struct Comm {
    ncclComm_t nccl;
    int n_ranks, rank;
};

}  // namespace dpr

using namespace dpr;

extern "C" {

int dpr_comm_unique_id(void* id128) {
    if (!id128) return DPR_ERR_NULL_POINTER;
    NcclApi& api = nccl_api();
    if (!api.ok) return nccl_fail(ncclSystemError, "dlopen(libnccl.so.2)");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    ncclResult_t r = api.GetUniqueId(&id);
    if (r != ncclSuccess) return nccl_fail(r, "ncclGetUniqueId");
    memcpy(id128, &id, sizeof(id));
    return DPR_OK;
}

int dpr_comm_init_rank(dpr_comm_t* comm, int n_ranks, int rank, const void* id128) {
    if (!comm || !id128) return DPR_ERR_NULL_POINTER;
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return DPR_ERR_BAD_DIMS;
    NcclApi& api = nccl_api();
    if (!api.ok) return nccl_fail(ncclSystemError, "dlopen(libnccl.so.2)");
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    Comm* c = new Comm{nullptr, n_ranks, rank};
    ncclResult_t r = api.CommInitRank(&c->nccl, n_ranks, id, rank);
    if (r != ncclSuccess) { delete c; return nccl_fail(r, "ncclCommInitRank"); }
    *comm = reinterpret_cast<dpr_comm_t>(c);
    return DPR_OK;
}

int dpr_comm_destroy(dpr_comm_t comm) {
    if (!comm) return DPR_OK;
    Comm* c = reinterpret_cast<Comm*>(comm);
    NcclApi& api = nccl_api();
    ncclResult_t r = api.ok ? api.CommDestroy(c->nccl) : ncclSuccess;
    delete c;
    return r == ncclSuccess ? DPR_OK : nccl_fail(r, "ncclCommDestroy");
}

static int allreduce(dpr_comm_t comm, void* buf, int64_t count, ncclDataType_t dt, dpr_stream_t stream) {
    if (!comm || (!buf && count > 0)) return DPR_ERR_NULL_POINTER;
    if (count < 0) return DPR_ERR_BAD_DIMS;
    if (count == 0) return DPR_OK;
    Comm* c = reinterpret_cast<Comm*>(comm);
    NcclApi& api = nccl_api();
    if (!api.ok) return nccl_fail(ncclSystemError, "dlopen(libnccl.so.2)");
    ncclResult_t r = api.AllReduce(buf, buf, (size_t)count, dt, ncclSum, c->nccl, static_cast<cudaStream_t>(stream));
    return r == ncclSuccess ? DPR_OK : nccl_fail(r, "ncclAllReduce");
}

int dpr_comm_allreduce_sum_f32(dpr_comm_t comm, float* buf, int64_t count, dpr_stream_t stream) {
    return allreduce(comm, buf, count, ncclFloat32, stream);
}
int dpr_comm_allreduce_sum_f64(dpr_comm_t comm, double* buf, int64_t count, dpr_stream_t stream) {
    return allreduce(comm, buf, count, ncclFloat64, stream);
}

}  // extern "C"
