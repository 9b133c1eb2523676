// dpr_comm.cu - the multi-GPU entry points of the C ABI (include/dpr.h): one process per GPU, pose-sharded batch,
// ONE all-reduce (sum) of the packed [d_points; d_point_weight] buffer per pullback (src/raster_pullback.jl:141,146).
//
// Two implementations behind the same entry points:
//   * one-shot peer-memory all-reduce (this file's kernels): every rank owns a symmetric buffer that its peers map through
//     CUDA IPC; a call copies the payload into it and raises a flag in every peer's memory over NVLink (kernel 1), waits
//     for the peers' flags and then sums all ranks' buffers with plain peer loads (kernel 2) - no ring, no intermediate
//     hops, the same summation order on every rank (bit-identical results across ranks).  Used for payloads up to
//     kP2pCapacity (config 2: 1.6 MB, config 5: 16 MB) when the IPC set-up succeeded.
//   * NCCL (resolved at run time with dlopen of libnccl.so.2 - the copy the host process already loaded, e.g. the one
//     bundled with PyTorch or NCCL.jl's artifact; the few NCCL types needed are declared here, so building libdpr.so
//     needs neither NCCL's headers nor its library): bootstrap (the IPC handles travel through ncclAllGather), larger
//     payloads, and the fallback when peer access is not available.
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <mutex>

#include "dpr_internal.h"

// ---- the part of NCCL's C ABI this file uses (stable across NCCL 2.x) --------------------------------------------
extern "C" {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0, ncclUnhandledCudaError = 1, ncclSystemError = 2, ncclInternalError = 3, ncclInvalidArgument = 4 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclFloat32 = 7, ncclFloat64 = 8 } ncclDataType_t;
typedef enum { ncclSum = 0 } ncclRedOp_t;
}

namespace dpr {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
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
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(api.handle, "ncclAllGather"));
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

// ---- one-shot all-reduce over peer memory ---------------------------------------------------------------------------
constexpr int kMaxRanks = 16;
constexpr size_t kP2pCapacity = (size_t)16 << 20;     // payload bytes per call served by the peer-memory path
constexpr size_t kFlagsBytes = 8192;                  // "published" flags: [rank] x 128-byte slots; CTA counters at 2048 and 2176;
constexpr size_t kCounterOffset = 2048;               // "slice reduced" flags (two-shot): 4096 + [rank] x 128
constexpr size_t kCounter2Offset = 2176;
constexpr size_t kFlagsBOffset = 4096;
// two-shot (reduce-scatter + all-gather) instead of one-shot once the one-shot's peer reads exceed this many bytes per rank
constexpr size_t kTwoShotPeerBytes = (size_t)32 << 20;

struct PeerTable {
    char* base[kMaxRanks];        // symmetric buffer of every rank as mapped into this process (base[rank] = own)
};

struct Comm {
    ncclComm_t nccl;
    int n_ranks, rank;
    bool p2p = false;
    char* sym = nullptr;                      // own symmetric buffer: flags | data (parity 0) | data (parity 1)
    PeerTable peers{};
    unsigned long long epoch = 0;
    int sm_count = 0;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Two launches on the caller's stream, deadlock-free by construction (no CTA ever waits for another CTA of its own grid):
//   publish: payload -> own symmetric buffer; the last CTA to finish raises this rank's flag in EVERY rank's flag array
//            (st.release.sys over NVLink).  Waits for nothing.
//   reduce:  every CTA waits until every rank's flag in the own array shows this epoch (ld.acquire.sys), then
//            buf[i] = sum over ranks, in rank order, of their buffers (plain peer loads).
// Buffers alternate with the epoch's parity: a rank that publishes epoch e + 2 has seen every peer's flag of epoch e + 1,
// which a peer raises only after its reduce kernel of epoch e - the last reader of that buffer - has finished.
template <typename T>
__global__ void __launch_bounds__(512) allreduce_publish_kernel(const T* __restrict__ buf, int64_t count, PeerTable peers, int n_ranks, int rank,
                                                                unsigned long long epoch) {
    constexpr int VEC = 16 / sizeof(T);
    const size_t data_off = kFlagsBytes + (size_t)(epoch & 1ull) * kP2pCapacity;
    const int64_t n_vec = count / VEC;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    T* mine = reinterpret_cast<T*>(peers.base[rank] + data_off);
    if ((reinterpret_cast<uintptr_t>(buf) % 16) == 0) {
        const uint4* src = reinterpret_cast<const uint4*>(buf);
        uint4* dst = reinterpret_cast<uint4*>(mine);
        for (int64_t i = first; i < n_vec; i += stride) dst[i] = src[i];
        for (int64_t i = n_vec * VEC + first; i < count; i += stride) mine[i] = buf[i];
    } else {
        for (int64_t i = first; i < count; i += stride) mine[i] = buf[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int* counter = reinterpret_cast<unsigned int*>(peers.base[rank] + kCounterOffset);
        if (atomicAdd(counter, 1u) == gridDim.x - 1) {          // this rank's payload is complete
            *counter = 0u;
            __threadfence_system();
            for (int r = 0; r < n_ranks; ++r) st_release_sys(reinterpret_cast<unsigned long long*>(peers.base[r] + (size_t)rank * 128), epoch);
        }
    }
}
template <typename T>
__global__ void __launch_bounds__(512) allreduce_reduce_kernel(T* __restrict__ buf, int64_t count, PeerTable peers, int n_ranks, int rank,
                                                               unsigned long long epoch) {
    constexpr int VEC = 16 / sizeof(T);
    const size_t data_off = kFlagsBytes + (size_t)(epoch & 1ull) * kP2pCapacity;
    const int64_t n_vec = count / VEC;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(buf) % 16) == 0;
    if (threadIdx.x < n_ranks) {
        const unsigned long long* f = reinterpret_cast<const unsigned long long*>(peers.base[rank] + (size_t)threadIdx.x * 128);
        while (ld_acquire_sys(f) < epoch) {}
    }
    __syncthreads();
    if (vec_ok) {
        struct alignas(16) Pack { T v[VEC]; };
        for (int64_t i = first; i < n_vec; i += stride) {
            Pack acc;
#pragma unroll
            for (int k = 0; k < VEC; ++k) acc.v[k] = T(0);
            for (int r = 0; r < n_ranks; ++r) {
                const Pack p = reinterpret_cast<const Pack*>(peers.base[r] + data_off)[i];
#pragma unroll
                for (int k = 0; k < VEC; ++k) acc.v[k] += p.v[k];
            }
            reinterpret_cast<Pack*>(buf)[i] = acc;
        }
    }
    for (int64_t i = (vec_ok ? n_vec * VEC : 0) + first; i < count; i += stride) {
        T acc = T(0);
        for (int r = 0; r < n_ranks; ++r) acc += reinterpret_cast<const T*>(peers.base[r] + data_off)[i];
        buf[i] = acc;
    }
}

// Two-shot variant for large payloads on many ranks (the one-shot kernel reads (n - 1) x payload over NVLink per rank:
// 112 MB for 16 MB on 8 ranks, 0.19 ms; two shots read 2 x (n - 1) / n x payload = 28 MB).  After the same publish kernel:
//   reduce-scatter: rank r sums slice r of every rank's buffer (rank order), stores it into slice r of its OWN symmetric
//                   buffer - no peer reads that slice before the second flag - and into buf; the last CTA raises the
//                   rank's "slice reduced" flag in every rank's flag array.
//   all-gather:     waits for every rank's second flag, copies slice q from rank q's buffer into buf.
// Every element is summed by exactly one rank, in rank order, so all ranks again hold bit-identical results.
template <typename T>
__device__ __forceinline__ void store_vec(T* buf, int64_t i_vec, const T (&v)[16 / sizeof(T)], bool vec_ok) {
    constexpr int VEC = 16 / sizeof(T);
    struct alignas(16) Pack { T v[VEC]; };
    if (vec_ok) {
        Pack p;
#pragma unroll
        for (int k = 0; k < VEC; ++k) p.v[k] = v[k];
        reinterpret_cast<Pack*>(buf)[i_vec] = p;
    } else {
#pragma unroll
        for (int k = 0; k < VEC; ++k) buf[i_vec * VEC + k] = v[k];
    }
}
template <typename T>
__global__ void __launch_bounds__(512) allreduce_reduce_scatter_kernel(T* __restrict__ buf, int64_t count, PeerTable peers, int n_ranks,
                                                                       int rank, unsigned long long epoch) {
    constexpr int VEC = 16 / sizeof(T);
    struct alignas(16) Pack { T v[VEC]; };
    const size_t data_off = kFlagsBytes + (size_t)(epoch & 1ull) * kP2pCapacity;
    const int64_t n_vec = count / VEC, per = (n_vec + n_ranks - 1) / n_ranks;
    const int64_t lo = (int64_t)rank * per, hi = lo + per < n_vec ? lo + per : n_vec;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(buf) % 16) == 0;
    if (threadIdx.x < n_ranks) {
        const unsigned long long* f = reinterpret_cast<const unsigned long long*>(peers.base[rank] + (size_t)threadIdx.x * 128);
        while (ld_acquire_sys(f) < epoch) {}
    }
    __syncthreads();
    Pack* own = reinterpret_cast<Pack*>(peers.base[rank] + data_off);
    for (int64_t i = lo + first; i < hi; i += stride) {
        T acc[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[k] = T(0);
        for (int r = 0; r < n_ranks; ++r) {
            const Pack p = reinterpret_cast<const Pack*>(peers.base[r] + data_off)[i];
#pragma unroll
            for (int k = 0; k < VEC; ++k) acc[k] += p.v[k];
        }
        Pack o;
#pragma unroll
        for (int k = 0; k < VEC; ++k) o.v[k] = acc[k];
        own[i] = o;
        store_vec<T>(buf, i, acc, vec_ok);
    }
    for (int64_t i = n_vec * VEC + first; i < count; i += stride) {          // fewer than VEC trailing elements: every rank sums them
        T acc = T(0);
        for (int r = 0; r < n_ranks; ++r) acc += reinterpret_cast<const T*>(peers.base[r] + data_off)[i];
        buf[i] = acc;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int* counter = reinterpret_cast<unsigned int*>(peers.base[rank] + kCounter2Offset);
        if (atomicAdd(counter, 1u) == gridDim.x - 1) {          // this rank's slice is complete
            *counter = 0u;
            __threadfence_system();
            for (int r = 0; r < n_ranks; ++r)
                st_release_sys(reinterpret_cast<unsigned long long*>(peers.base[r] + kFlagsBOffset + (size_t)rank * 128), epoch);
        }
    }
}
template <typename T>
__global__ void __launch_bounds__(512) allreduce_all_gather_kernel(T* __restrict__ buf, int64_t count, PeerTable peers, int n_ranks, int rank,
                                                                   unsigned long long epoch) {
    constexpr int VEC = 16 / sizeof(T);
    struct alignas(16) Pack { T v[VEC]; };
    const size_t data_off = kFlagsBytes + (size_t)(epoch & 1ull) * kP2pCapacity;
    const int64_t n_vec = count / VEC, per = (n_vec + n_ranks - 1) / n_ranks;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(buf) % 16) == 0;
    if (threadIdx.x < n_ranks) {
        const unsigned long long* f = reinterpret_cast<const unsigned long long*>(peers.base[rank] + kFlagsBOffset + (size_t)threadIdx.x * 128);
        while (ld_acquire_sys(f) < epoch) {}
    }
    __syncthreads();
    for (int q = 0; q < n_ranks; ++q) {
        if (q == rank) continue;
        const int64_t lo = (int64_t)q * per, hi = lo + per < n_vec ? lo + per : n_vec;
        const Pack* __restrict__ src = reinterpret_cast<const Pack*>(peers.base[q] + data_off);
        for (int64_t i = lo + first; i < hi; i += stride) {
            const Pack p = src[i];
            T v[VEC];
#pragma unroll
            for (int k = 0; k < VEC; ++k) v[k] = p.v[k];
            store_vec<T>(buf, i, v, vec_ok);
        }
    }
}

// Symmetric buffers + IPC handle exchange (through NCCL, on the current device).  Any failure leaves the NCCL path.
static void p2p_setup(Comm* c) {
    NcclApi& api = nccl_api();
    if (c->n_ranks < 2 || c->n_ranks > kMaxRanks || !api.AllGather) return;
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, dev);
    const size_t bytes = kFlagsBytes + 2 * kP2pCapacity;
    bool ok = cudaMalloc(&c->sym, bytes) == cudaSuccess && cudaMemset(c->sym, 0, kFlagsBytes) == cudaSuccess;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    ok = ok && cudaIpcGetMemHandle(&mine, c->sym) == cudaSuccess;
    // every rank takes part in the exchange, also one whose allocation failed (it sends a zero handle and a 0 flag)
    struct Msg { cudaIpcMemHandle_t h; int ok; int pad[15]; };
    static_assert(sizeof(Msg) == sizeof(cudaIpcMemHandle_t) + 64, "message layout");
    Msg msg;
    memset(&msg, 0, sizeof(msg));
    msg.h = mine;
    msg.ok = ok ? 1 : 0;
    Msg* d_msgs = nullptr;
    Msg* h_msgs = new Msg[c->n_ranks];
    cudaStream_t s = nullptr;
    bool xok = cudaMalloc(&d_msgs, sizeof(Msg) * (size_t)(c->n_ranks + 1)) == cudaSuccess && cudaStreamCreate(&s) == cudaSuccess;
    if (xok) {
        xok = cudaMemcpyAsync(d_msgs + c->n_ranks, &msg, sizeof(Msg), cudaMemcpyHostToDevice, s) == cudaSuccess &&
              api.AllGather(d_msgs + c->n_ranks, d_msgs, sizeof(Msg), ncclInt8, c->nccl, s) == ncclSuccess &&
              cudaMemcpyAsync(h_msgs, d_msgs, sizeof(Msg) * (size_t)c->n_ranks, cudaMemcpyDeviceToHost, s) == cudaSuccess &&
              cudaStreamSynchronize(s) == cudaSuccess;
    }
    bool all = xok;
    for (int r = 0; all && r < c->n_ranks; ++r) all = h_msgs[r].ok == 1;
    for (int r = 0; all && r < c->n_ranks; ++r) {
        if (r == c->rank) { c->peers.base[r] = c->sym; continue; }
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, h_msgs[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { all = false; break; }
        c->peers.base[r] = static_cast<char*>(p);
    }
    // agree on the outcome: a rank that could not map a peer must not leave the others spinning in the kernel
    if (xok) {
        int flag = all ? 1 : 0;
        int* d_flags = reinterpret_cast<int*>(d_msgs);
        int h_flags[kMaxRanks + 1];
        xok = cudaMemcpyAsync(d_flags + c->n_ranks, &flag, sizeof(int), cudaMemcpyHostToDevice, s) == cudaSuccess &&
              api.AllGather(d_flags + c->n_ranks, d_flags, sizeof(int), ncclInt8, c->nccl, s) == ncclSuccess &&
              cudaMemcpyAsync(h_flags, d_flags, sizeof(int) * (size_t)c->n_ranks, cudaMemcpyDeviceToHost, s) == cudaSuccess &&
              cudaStreamSynchronize(s) == cudaSuccess;
        for (int r = 0; xok && r < c->n_ranks; ++r) all = all && h_flags[r] == 1;
    }
    c->p2p = xok && all;
    cudaGetLastError();
    if (s) cudaStreamDestroy(s);
    if (d_msgs) cudaFree(d_msgs);
    delete[] h_msgs;
}

static void p2p_teardown(Comm* c) {
    for (int r = 0; r < c->n_ranks && r < kMaxRanks; ++r)
        if (r != c->rank && c->peers.base[r]) cudaIpcCloseMemHandle(c->peers.base[r]);
    if (c->sym) cudaFree(c->sym);
    c->sym = nullptr;
    c->p2p = false;
}

}  // namespace dpr

using namespace dpr;

template <typename T>
static int allreduce(dpr_comm_t comm, T* buf, int64_t count, ncclDataType_t dt, dpr_stream_t stream) {
    if (!comm || (!buf && count > 0)) return DPR_ERR_NULL_POINTER;
    if (count < 0) return DPR_ERR_BAD_DIMS;
    if (count == 0) return DPR_OK;
    Comm* c = reinterpret_cast<Comm*>(comm);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (c->p2p && tuning().comm_p2p != 1 && (size_t)count * sizeof(T) <= kP2pCapacity) {
        ++c->epoch;                                         // every rank calls in the same order (collective semantics)
        int64_t ctas = (count * (int64_t)sizeof(T) / 16 + 511) / 512;
        if (ctas > (int64_t)c->sm_count * 2) ctas = (int64_t)c->sm_count * 2;
        if (ctas < 1) ctas = 1;
        {
            LaunchScope scope("allreduce_p2p_publish", s);
            allreduce_publish_kernel<T><<<(unsigned)ctas, 512, 0, s>>>(buf, count, c->peers, c->n_ranks, c->rank, c->epoch);
        }
        // (the choice depends on the payload size and the rank count only: every rank takes the same path)
        const bool two_shot = c->n_ranks >= 4 && (size_t)count * sizeof(T) * (size_t)(c->n_ranks - 1) >= kTwoShotPeerBytes;
        if (two_shot) {
            {
                LaunchScope scope("allreduce_p2p_reduce_scatter", s);
                allreduce_reduce_scatter_kernel<T><<<(unsigned)ctas, 512, 0, s>>>(buf, count, c->peers, c->n_ranks, c->rank, c->epoch);
            }
            {
                LaunchScope scope("allreduce_p2p_all_gather", s);
                allreduce_all_gather_kernel<T><<<(unsigned)ctas, 512, 0, s>>>(buf, count, c->peers, c->n_ranks, c->rank, c->epoch);
            }
        } else {
            LaunchScope scope("allreduce_p2p_reduce", s);
            allreduce_reduce_kernel<T><<<(unsigned)ctas, 512, 0, s>>>(buf, count, c->peers, c->n_ranks, c->rank, c->epoch);
        }
        DPR_CUDA_TRY(cudaGetLastError());
        return DPR_OK;
    }
    NcclApi& api = nccl_api();
    if (!api.ok) return nccl_fail(ncclSystemError, "dlopen(libnccl.so.2)");
    ncclResult_t r = api.AllReduce(buf, buf, (size_t)count, dt, ncclSum, c->nccl, s);
    return r == ncclSuccess ? DPR_OK : nccl_fail(r, "ncclAllReduce");
}

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
    Comm* c = new Comm();
    c->nccl = nullptr; c->n_ranks = n_ranks; c->rank = rank;
    ncclResult_t r = api.CommInitRank(&c->nccl, n_ranks, id, rank);
    if (r != ncclSuccess) { delete c; return nccl_fail(r, "ncclCommInitRank"); }
    if (tuning().comm_p2p != 1) p2p_setup(c);
    *comm = reinterpret_cast<dpr_comm_t>(c);
    return DPR_OK;
}

int dpr_comm_destroy(dpr_comm_t comm) {
    if (!comm) return DPR_OK;
    Comm* c = reinterpret_cast<Comm*>(comm);
    NcclApi& api = nccl_api();
    p2p_teardown(c);
    ncclResult_t r = api.ok ? api.CommDestroy(c->nccl) : ncclSuccess;
    delete c;
    return r == ncclSuccess ? DPR_OK : nccl_fail(r, "ncclCommDestroy");
}

/* 1 when the communicator serves small payloads with the one-shot peer-memory kernel, 0 when everything goes through NCCL */
int dpr_comm_uses_peer_memory(dpr_comm_t comm) { return comm && reinterpret_cast<Comm*>(comm)->p2p ? 1 : 0; }

int dpr_comm_allreduce_sum_f32(dpr_comm_t comm, float* buf, int64_t count, dpr_stream_t stream) {
    return allreduce<float>(comm, buf, count, ncclFloat32, stream);
}
int dpr_comm_allreduce_sum_f64(dpr_comm_t comm, double* buf, int64_t count, dpr_stream_t stream) {
    return allreduce<double>(comm, buf, count, ncclFloat64, stream);
}

}  // extern "C"
