// dpr_api.cu - the C ABI of libdpr.so (include/dpr.h): validation, dispatch, options, host-buffer entry points.
// No torch types, no C++ types in the signatures; never throws; there is NO CPU fallback.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <thread>
#include <utility>
#include <vector>

#include "dpr_internal.h"

namespace dpr {

static Tuning g_tuning;
static std::atomic<int64_t> g_launches{0};
static thread_local char tl_error[512] = "";
static thread_local const char* tl_path[2] = {"none", "none"};

const Tuning& tuning() { return g_tuning; }
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
void set_last_path(int op, const char* name) { if (op >= 0 && op < 2) tl_path[op] = name; }

// ---- optional per-kernel event timing -----------------------------------------------------------------
struct ProfileRecord { const char* name; cudaEvent_t e0, e1; };
static std::mutex g_prof_mutex;
static std::vector<ProfileRecord> g_prof;
static std::atomic<int> g_prof_on{0};

LaunchScope::LaunchScope(const char* n, cudaStream_t s) : name(n), stream(s), slot(-1) {
    count_launches(1);
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return;
    cudaEventRecord(e0, s);
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    g_prof.emplace_back(ProfileRecord{n, e0, e1});
    slot = (int)g_prof.size() - 1;
}
LaunchScope::~LaunchScope() {
    if (g_prof_on.load(std::memory_order_relaxed) == 2) {
        // checked mode (dpr_profile_enable(2)): wait for the kernel and name it if it faulted
        const cudaError_t e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) {
            char buf[400];
            snprintf(buf, sizeof(buf), "kernel %s failed: %s (%s)", name, cudaGetErrorName(e), cudaGetErrorString(e));
            set_error_message(buf);
            fprintf(stderr, "libdpr: %s\n", buf);
        }
    }
    if (slot < 0) return;
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    if (slot < (int)g_prof.size()) cudaEventRecord(g_prof[slot].e1, stream);
}
static void profile_clear() {
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    for (auto& r : g_prof) { if (r.e0) cudaEventDestroy(r.e0); if (r.e1) cudaEventDestroy(r.e1); }
    g_prof.clear();
}

void set_error_message(const char* msg) { snprintf(tl_error, sizeof(tl_error), "%s", msg); }

int cuda_fail(cudaError_t e, const char* what) {
    snprintf(tl_error, sizeof(tl_error), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    cudaGetLastError();  // clear the sticky-free error state
    return DPR_ERR_CUDA;
}

// largest dynamic shared-memory size granted so far per (kernel, device)
static std::mutex g_smem_mutex;
static std::map<std::pair<const void*, int>, size_t> g_smem_granted;
int opt_in_smem(const void* kernel, size_t bytes, int device) {
    if (bytes <= 48 * 1024) return DPR_OK;          // available without opting in
    std::lock_guard<std::mutex> lock(g_smem_mutex);
    size_t& have = g_smem_granted[std::make_pair(kernel, device)];
    if (bytes <= have) return DPR_OK;
    DPR_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    have = bytes;
    return DPR_OK;
}

// per-device attribute cache (the only global mutable state besides options and the host staging arenas)
static std::mutex g_dev_mutex;
static DeviceInfo g_dev_cache[64];

int current_device_info(DeviceInfo& info) {
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { cuda_fail(e, "cudaGetDevice"); return DPR_ERR_NO_DEVICE; }
    if (dev < 0 || dev >= 64) return DPR_ERR_NO_DEVICE;
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    DeviceInfo& c = g_dev_cache[dev];
    if (c.device != dev) {
        DeviceInfo d;
        d.device = dev;
        DPR_CUDA_TRY(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev));
        DPR_CUDA_TRY(cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
        DPR_CUDA_TRY(cudaDeviceGetAttribute(&d.cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
        DPR_CUDA_TRY(cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        if (d.cc_major != 10) {
            snprintf(tl_error, sizeof(tl_error), "device %d is sm_%d%d; libdpr.so holds sm_100a kernels only", dev,
                     d.cc_major, d.cc_minor);
            return DPR_ERR_NO_DEVICE;
        }
        c = d;
    }
    info = c;
    return DPR_OK;
}

static int check_dims(int n_in, int n_out, const int64_t* grid, int64_t P, int64_t B) {
    if (!grid) return DPR_ERR_NULL_POINTER;
    if (n_in < 1 || n_in > kMaxDim || n_out < 1 || n_out > kMaxDim) return DPR_ERR_UNSUPPORTED;   // any 1 <= N_in, N_out <= 4
    if (P < 0 || B < 0) return DPR_ERR_BAD_DIMS;
    int64_t cells = 1;
    for (int k = 0; k < n_out; ++k) {
        if (grid[k] < 1 || grid[k] > (int64_t)1 << 20) return DPR_ERR_BAD_DIMS;
        cells *= grid[k];
        if (cells > (int64_t)1 << 40) return DPR_ERR_BAD_DIMS;
    }
    if (B > 0 && cells > ((int64_t)1 << 46) / B) return DPR_ERR_BAD_DIMS;
    if (P > (int64_t)1 << 40) return DPR_ERR_BAD_DIMS;
    return DPR_OK;
}

template <typename T>
static int forward_entry(int n_in, int n_out, const int64_t* grid, int64_t P, int64_t B, const T* points,
                         const T* rotation, const T* translation, const T* background, const T* out_weight,
                         const T* point_weight, T* out, void* workspace, size_t workspace_bytes, dpr_stream_t stream) {
    int rc = check_dims(n_in, n_out, grid, P, B);
    if (rc != DPR_OK) return rc;
    if ((P > 0 && !points) || (B > 0 && (!rotation || !translation || !out))) return DPR_ERR_NULL_POINTER;
    DeviceInfo dev;
    rc = current_device_info(dev);
    if (rc != DPR_OK) return rc;
    if (!workspace || workspace_bytes < 256) return DPR_ERR_WORKSPACE;   // smaller than dpr_workspace_bytes(): no point sort
    ForwardArgs<T> a;
    a.n_in = n_in; a.n_out = n_out;
    for (int k = 0; k < kMaxDim; ++k) a.grid[k] = k < n_out ? grid[k] : 1;
    a.P = P; a.B = B;
    a.points = points; a.rotation = rotation; a.translation = translation;
    a.background = background; a.out_weight = out_weight; a.point_weight = point_weight;
    a.out = out; a.workspace = workspace; a.workspace_bytes = workspace_bytes;
    a.stream = static_cast<cudaStream_t>(stream);
    if (B == 0) return DPR_OK;
    return forward_dispatch<T>(a, dev);
}

template <typename T>
static int pullback_entry(int n_in, int n_out, const int64_t* grid, int64_t P, int64_t B, const T* ds_dout,
                          const T* points, const T* rotation, const T* translation, const T* out_weight,
                          const T* point_weight, T* d_points, T* d_rotation, T* d_translation, T* d_background,
                          T* d_out_weight, T* d_point_weight, void* workspace, size_t workspace_bytes,
                          dpr_stream_t stream) {
    int rc = check_dims(n_in, n_out, grid, P, B);
    if (rc != DPR_OK) return rc;
    if ((P > 0 && (!points || !d_points)) || (B > 0 && (!rotation || !translation || !ds_dout || !d_rotation || !d_translation)))
        return DPR_ERR_NULL_POINTER;
    DeviceInfo dev;
    rc = current_device_info(dev);
    if (rc != DPR_OK) return rc;
    if (workspace_bytes < 256) return DPR_ERR_WORKSPACE;   // smaller than dpr_workspace_bytes(): the point sort is skipped
    PullbackArgs<T> a;
    a.n_in = n_in; a.n_out = n_out;
    for (int k = 0; k < kMaxDim; ++k) a.grid[k] = k < n_out ? grid[k] : 1;
    a.P = P; a.B = B;
    a.ds_dout = ds_dout; a.points = points; a.rotation = rotation; a.translation = translation;
    a.out_weight = out_weight; a.point_weight = point_weight;
    a.d_points = d_points; a.d_rotation = d_rotation; a.d_translation = d_translation;
    a.d_background = d_background; a.d_out_weight = d_out_weight; a.d_point_weight = d_point_weight;
    a.workspace = workspace; a.workspace_bytes = workspace_bytes;
    a.stream = static_cast<cudaStream_t>(stream);
    return pullback_dispatch<T>(a, dev);
}

// ---------------------------------------------------------------------------------------------------------
// Host-buffer entry points: pose chunks are pipelined through a per-device staging arena on NSTREAM streams so
// that the H2D copy of chunk i+1, the kernels of chunk i and the D2H copy of chunk i-1 overlap.
// ---------------------------------------------------------------------------------------------------------
static constexpr int NSTREAM = 3;

struct HostArena {
    int device = -1;
    cudaStream_t streams[NSTREAM] = {nullptr, nullptr, nullptr};
    cudaEvent_t shared_ready = nullptr;
    void* shared = nullptr;   size_t shared_bytes = 0;    // points, point_weight, pose-summed gradients
    void* slot[NSTREAM] = {nullptr, nullptr, nullptr};
    size_t slot_bytes = 0;
    void* ws[NSTREAM] = {nullptr, nullptr, nullptr};   // per-stream kernel workspace (dpr_workspace_bytes)
    size_t ws_bytes = 0;
};
// One arena (and one lock) per device AND operation: a process driving several GPUs from several threads is not serialised,
// and a forward and a pullback on the same device can run at the same time - their copies then use both directions of the
// host link (dpr_raster_*_host_async_*).
static std::mutex g_arena_mutex[64][2];
static HostArena g_arena[64][2];

static int arena_device(int& dev) {
    dev = -1;
    DPR_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return DPR_ERR_NO_DEVICE;
    return DPR_OK;
}
// caller holds g_arena_mutex[dev][op]
static int arena_get(int dev, int op, HostArena*& out) {
    HostArena& a = g_arena[dev][op];
    if (a.device != dev) {
        // create everything or nothing: a failed call must not leave half an arena behind (the next call would create
        // the streams again and leak the first set)
        cudaError_t e = cudaSuccess;
        for (int i = 0; i < NSTREAM && e == cudaSuccess; ++i) e = cudaStreamCreateWithFlags(&a.streams[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&a.shared_ready, cudaEventDisableTiming);
        for (int i = 0; i < NSTREAM && e == cudaSuccess; ++i) e = cudaMalloc(&a.ws[i], 4096);
        for (int i = 0; i < NSTREAM && e == cudaSuccess; ++i) e = cudaMemset(a.ws[i], 0, 256);      // clean binning-cache header
        if (e != cudaSuccess) {
            for (int i = 0; i < NSTREAM; ++i) {
                if (a.streams[i]) cudaStreamDestroy(a.streams[i]);
                if (a.ws[i]) cudaFree(a.ws[i]);
                a.streams[i] = nullptr; a.ws[i] = nullptr;
            }
            if (a.shared_ready) cudaEventDestroy(a.shared_ready);
            a.shared_ready = nullptr;
            return cuda_fail(e, "host arena setup");
        }
        a.ws_bytes = 4096;
        a.device = dev;
    }
    out = &a;
    return DPR_OK;
}
static int arena_reserve(void*& ptr, size_t& have, size_t want) {
    if (want <= have) return DPR_OK;
    if (ptr) DPR_CUDA_TRY(cudaFree(ptr));
    ptr = nullptr; have = 0;
    want = (want + 4095) / 4096 * 4096;
    DPR_CUDA_TRY(cudaMalloc(&ptr, want));
    DPR_CUDA_TRY(cudaMemset(ptr, 0, 256));      // a kernel workspace starts with a clean binning-cache header (include/dpr.h)
    have = want;
    return DPR_OK;
}
static size_t align256(size_t n) { return (n + 255) / 256 * 256; }

// poses per chunk: ~64 MB of image data per slot keeps copies long enough to reach PCIe rate
static int64_t host_chunk_poses(int64_t cells, int64_t B, size_t sizeof_T) {
    const int64_t target = (int64_t)64 << 20;
    int64_t n = target / (cells * (int64_t)sizeof_T);
    if (n < 1) n = 1;
    if (n > B) n = B;
    return n;
}

template <typename T>
static int forward_host(int n_in, int n_out, const int64_t* grid, int64_t P, int64_t B, const T* points,
                        const T* rotation, const T* translation, const T* background, const T* out_weight,
                        const T* point_weight, T* out) {
    int rc = check_dims(n_in, n_out, grid, P, B);
    if (rc != DPR_OK) return rc;
    if ((P > 0 && !points) || (B > 0 && (!rotation || !translation || !out))) return DPR_ERR_NULL_POINTER;
    if (B == 0) return DPR_OK;
    int dev_id = -1;
    rc = arena_device(dev_id);
    if (rc != DPR_OK) return rc;
    std::lock_guard<std::mutex> lock(g_arena_mutex[dev_id][DPR_OP_FORWARD]);
    HostArena* ar = nullptr;
    rc = arena_get(dev_id, DPR_OP_FORWARD, ar);
    if (rc != DPR_OK) return rc;
    int64_t cells = 1;
    for (int k = 0; k < n_out; ++k) cells *= grid[k];
    const int64_t cb = host_chunk_poses(cells, B, sizeof(T));
    // The per-pose vectors of the WHOLE batch are copied once, up front (a few hundred KB): small copies issued per chunk
    // would queue behind a concurrent call's 64 MB copies in the same direction of the host link and stall this pipeline.
    const size_t pts_bytes = align256(sizeof(T) * (size_t)(P * n_in)), pw_bytes = align256(sizeof(T) * (size_t)P);
    const size_t rot_b = align256(sizeof(T) * (size_t)(B * n_out * n_in)), tr_b = align256(sizeof(T) * (size_t)(B * n_out));
    const size_t vec_b = align256(sizeof(T) * (size_t)B), img_b = align256(sizeof(T) * (size_t)(cb * cells));
    rc = arena_reserve(ar->shared, ar->shared_bytes, pts_bytes + pw_bytes + rot_b + tr_b + 2 * vec_b + 256);
    if (rc != DPR_OK) return rc;
    const size_t need = img_b;
    if (need > ar->slot_bytes) {
        size_t have = 0;
        for (int i = 0; i < NSTREAM; ++i) { have = ar->slot_bytes; rc = arena_reserve(ar->slot[i], have, need); if (rc != DPR_OK) return rc; }
        ar->slot_bytes = have;
    }
    {   // per-stream kernel workspace (point statistics, spatially sorted copy of the points)
        const size_t need_ws = forward_workspace_bytes(n_in, n_out, grid, P, cb, (int)sizeof(T));
        if (need_ws > ar->ws_bytes) {
            size_t have = 0;
            for (int i = 0; i < NSTREAM; ++i) { have = ar->ws_bytes; rc = arena_reserve(ar->ws[i], have, need_ws); if (rc != DPR_OK) return rc; }
            ar->ws_bytes = have;
        }
    }
    char* sh = static_cast<char*>(ar->shared);
    T* d_points = reinterpret_cast<T*>(sh);
    T* d_pw = reinterpret_cast<T*>(sh + pts_bytes);
    T* d_rot = reinterpret_cast<T*>(sh + pts_bytes + pw_bytes);
    T* d_tr = reinterpret_cast<T*>(sh + pts_bytes + pw_bytes + rot_b);
    T* d_bg = reinterpret_cast<T*>(sh + pts_bytes + pw_bytes + rot_b + tr_b);
    T* d_ow = reinterpret_cast<T*>(sh + pts_bytes + pw_bytes + rot_b + tr_b + vec_b);
    cudaStream_t s0 = ar->streams[0];
    // never return (DPR_CUDA_TRY) with copies into the caller's buffers still in flight
    struct Drain {
        HostArena* ar;
        ~Drain() { for (int i = 0; i < NSTREAM; ++i) cudaStreamSynchronize(ar->streams[i]); }
    } drain{ar};
    if (P > 0) DPR_CUDA_TRY(cudaMemcpyAsync(d_points, points, sizeof(T) * (size_t)(P * n_in), cudaMemcpyHostToDevice, s0));
    if (point_weight && P > 0) DPR_CUDA_TRY(cudaMemcpyAsync(d_pw, point_weight, sizeof(T) * (size_t)P, cudaMemcpyHostToDevice, s0));
    DPR_CUDA_TRY(cudaMemcpyAsync(d_rot, rotation, sizeof(T) * (size_t)(B * n_out * n_in), cudaMemcpyHostToDevice, s0));
    DPR_CUDA_TRY(cudaMemcpyAsync(d_tr, translation, sizeof(T) * (size_t)(B * n_out), cudaMemcpyHostToDevice, s0));
    if (background) DPR_CUDA_TRY(cudaMemcpyAsync(d_bg, background, sizeof(T) * (size_t)B, cudaMemcpyHostToDevice, s0));
    if (out_weight) DPR_CUDA_TRY(cudaMemcpyAsync(d_ow, out_weight, sizeof(T) * (size_t)B, cudaMemcpyHostToDevice, s0));
    DPR_CUDA_TRY(cudaEventRecord(ar->shared_ready, s0));
    int64_t chunk = 0;
    for (int64_t b0 = 0; b0 < B; b0 += cb, ++chunk) {
        const int64_t nb = (b0 + cb < B) ? cb : B - b0;
        const int si = (int)(chunk % NSTREAM);
        cudaStream_t st = ar->streams[si];
        T* d_out = static_cast<T*>(ar->slot[si]);
        DPR_CUDA_TRY(cudaStreamWaitEvent(st, ar->shared_ready, 0));
        rc = forward_entry<T>(n_in, n_out, grid, P, nb, d_points, d_rot + b0 * n_out * n_in, d_tr + b0 * n_out,
                              background ? d_bg + b0 : nullptr, out_weight ? d_ow + b0 : nullptr, point_weight ? d_pw : nullptr,
                              d_out, ar->ws[si], ar->ws_bytes, st);
        if (rc != DPR_OK) break;
        DPR_CUDA_TRY(cudaMemcpyAsync(out + b0 * cells, d_out, sizeof(T) * (size_t)(nb * cells), cudaMemcpyDeviceToHost, st));
    }
    for (int i = 0; i < NSTREAM; ++i) {
        cudaError_t e = cudaStreamSynchronize(ar->streams[i]);
        if (e != cudaSuccess && rc == DPR_OK) rc = cuda_fail(e, "cudaStreamSynchronize");
    }
    return rc;
}

template <typename T>
__global__ void __launch_bounds__(256) accumulate_kernel(T* __restrict__ dst, const T* __restrict__ src, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] += src[i];
}

template <typename T>
static int pullback_host(int n_in, int n_out, const int64_t* grid, int64_t P, int64_t B, const T* ds_dout,
                         const T* points, const T* rotation, const T* translation, const T* out_weight,
                         const T* point_weight, T* h_dp, T* h_drot, T* h_dtr, T* h_dbg, T* h_dow, T* h_dpw) {
    int rc = check_dims(n_in, n_out, grid, P, B);
    if (rc != DPR_OK) return rc;
    if ((P > 0 && (!points || !h_dp)) || (B > 0 && (!rotation || !translation || !ds_dout || !h_drot || !h_dtr)))
        return DPR_ERR_NULL_POINTER;
    int dev_id = -1;
    rc = arena_device(dev_id);
    if (rc != DPR_OK) return rc;
    std::lock_guard<std::mutex> lock(g_arena_mutex[dev_id][DPR_OP_PULLBACK]);
    HostArena* ar = nullptr;
    rc = arena_get(dev_id, DPR_OP_PULLBACK, ar);
    if (rc != DPR_OK) return rc;
    int64_t cells = 1;
    for (int k = 0; k < n_out; ++k) cells *= grid[k];
    const int64_t cb = B > 0 ? host_chunk_poses(cells, B, sizeof(T)) : 1;
    // shared: points, point_weight, total d_points (+d_point_weight), one partial buffer per stream, and the per-pose
    // vectors of the WHOLE batch in both directions: they cross the host link once (inputs up front, gradients at the end)
    // instead of chunk by chunk, where they would queue behind a concurrent call's 64 MB copies.
    const size_t pts_bytes = align256(sizeof(T) * (size_t)(P * n_in)), pw_bytes = align256(sizeof(T) * (size_t)P);
    const size_t grad_bytes = pts_bytes + pw_bytes;
    const size_t rot_b = align256(sizeof(T) * (size_t)(B * n_out * n_in)), tr_b = align256(sizeof(T) * (size_t)(B * n_out));
    const size_t vec_b = align256(sizeof(T) * (size_t)B), img_b = align256(sizeof(T) * (size_t)(cb * cells));
    const size_t pose_b = 2 * rot_b + 2 * tr_b + 3 * vec_b;
    rc = arena_reserve(ar->shared, ar->shared_bytes, pts_bytes + pw_bytes + grad_bytes * (1 + NSTREAM) + pose_b + 256);
    if (rc != DPR_OK) return rc;
    const size_t need = img_b;
    if (need > ar->slot_bytes) {
        size_t have = 0;
        for (int i = 0; i < NSTREAM; ++i) { have = ar->slot_bytes; rc = arena_reserve(ar->slot[i], have, need); if (rc != DPR_OK) return rc; }
        ar->slot_bytes = have;
    }
    {   // per-stream kernel workspace (holds the spatially sorted copy of the points)
        const size_t need_ws = pullback_workspace_bytes(n_in, n_out, grid, P, cb, (int)sizeof(T));
        if (need_ws > ar->ws_bytes) {
            size_t have = 0;
            for (int i = 0; i < NSTREAM; ++i) { have = ar->ws_bytes; rc = arena_reserve(ar->ws[i], have, need_ws); if (rc != DPR_OK) return rc; }
            ar->ws_bytes = have;
        }
    }
    char* sh = static_cast<char*>(ar->shared);
    T* d_points = reinterpret_cast<T*>(sh);
    T* d_pw = reinterpret_cast<T*>(sh + pts_bytes);
    T* tot_dp = reinterpret_cast<T*>(sh + pts_bytes + pw_bytes);
    T* tot_dpw = reinterpret_cast<T*>(sh + pts_bytes + pw_bytes + pts_bytes);
    char* pose = sh + pts_bytes + pw_bytes + grad_bytes * (1 + NSTREAM);
    T* d_rot = reinterpret_cast<T*>(pose);
    T* d_tr = reinterpret_cast<T*>(pose + rot_b);
    T* d_ow = reinterpret_cast<T*>(pose + rot_b + tr_b);
    T* g_rot = reinterpret_cast<T*>(pose + rot_b + tr_b + vec_b);
    T* g_tr = reinterpret_cast<T*>(pose + 2 * rot_b + tr_b + vec_b);
    T* g_bg = reinterpret_cast<T*>(pose + 2 * rot_b + 2 * tr_b + vec_b);
    T* g_ow = reinterpret_cast<T*>(pose + 2 * rot_b + 2 * tr_b + 2 * vec_b);
    cudaStream_t s0 = ar->streams[0];
    // events are destroyed and the streams drained on every exit path, including the early returns of DPR_CUDA_TRY
    struct Cleanup {
        HostArena* ar;
        cudaEvent_t done[NSTREAM] = {nullptr, nullptr, nullptr};
        ~Cleanup() {
            for (int i = 0; i < NSTREAM; ++i) cudaStreamSynchronize(ar->streams[i]);
            for (int i = 0; i < NSTREAM; ++i) if (done[i]) cudaEventDestroy(done[i]);
        }
    } cleanup{ar};
    cudaEvent_t (&done)[NSTREAM] = cleanup.done;
    if (P > 0) DPR_CUDA_TRY(cudaMemcpyAsync(d_points, points, sizeof(T) * (size_t)(P * n_in), cudaMemcpyHostToDevice, s0));
    if (point_weight && P > 0) DPR_CUDA_TRY(cudaMemcpyAsync(d_pw, point_weight, sizeof(T) * (size_t)P, cudaMemcpyHostToDevice, s0));
    if (B > 0) {
        DPR_CUDA_TRY(cudaMemcpyAsync(d_rot, rotation, sizeof(T) * (size_t)(B * n_out * n_in), cudaMemcpyHostToDevice, s0));
        DPR_CUDA_TRY(cudaMemcpyAsync(d_tr, translation, sizeof(T) * (size_t)(B * n_out), cudaMemcpyHostToDevice, s0));
        if (out_weight) DPR_CUDA_TRY(cudaMemcpyAsync(d_ow, out_weight, sizeof(T) * (size_t)B, cudaMemcpyHostToDevice, s0));
    }
    DPR_CUDA_TRY(cudaMemsetAsync(tot_dp, 0, grad_bytes, s0));
    DPR_CUDA_TRY(cudaEventRecord(ar->shared_ready, s0));
    int64_t chunk = 0;
    for (int64_t b0 = 0; b0 < B && rc == DPR_OK; b0 += cb, ++chunk) {
        const int64_t nb = (b0 + cb < B) ? cb : B - b0;
        const int si = (int)(chunk % NSTREAM);
        cudaStream_t st = ar->streams[si];
        T* d_img = static_cast<T*>(ar->slot[si]);
        T* part_dp = reinterpret_cast<T*>(sh + pts_bytes + pw_bytes + grad_bytes * (1 + si));
        T* part_dpw = reinterpret_cast<T*>(reinterpret_cast<char*>(part_dp) + pts_bytes);
        DPR_CUDA_TRY(cudaStreamWaitEvent(st, ar->shared_ready, 0));
        DPR_CUDA_TRY(cudaMemcpyAsync(d_img, ds_dout + b0 * cells, sizeof(T) * (size_t)(nb * cells), cudaMemcpyHostToDevice, st));
        rc = pullback_entry<T>(n_in, n_out, grid, P, nb, d_img, d_points, d_rot + b0 * n_out * n_in, d_tr + b0 * n_out,
                               out_weight ? d_ow + b0 : nullptr, point_weight ? d_pw : nullptr, part_dp, g_rot + b0 * n_out * n_in,
                               g_tr + b0 * n_out, h_dbg ? g_bg + b0 : nullptr, h_dow ? g_ow + b0 : nullptr,
                               h_dpw ? part_dpw : nullptr, ar->ws[si], ar->ws_bytes, st);
        if (rc != DPR_OK) break;
        // fold this chunk's pose-sum into the total on stream 0 (serialises the adds, keeps them race-free)
        if (!done[si]) DPR_CUDA_TRY(cudaEventCreateWithFlags(&done[si], cudaEventDisableTiming));
        DPR_CUDA_TRY(cudaEventRecord(done[si], st));
        DPR_CUDA_TRY(cudaStreamWaitEvent(s0, done[si], 0));
        const int64_t n_acc = (int64_t)(grad_bytes / sizeof(T));
        int64_t blocks = (n_acc + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        if (n_acc > 0) {
            LaunchScope scope("accumulate", s0);
            accumulate_kernel<T><<<(unsigned)blocks, 256, 0, s0>>>(tot_dp, part_dp, n_acc);
        }
        // the slot's partial buffer may only be reused after the fold: make the slot's stream wait for it
        DPR_CUDA_TRY(cudaEventRecord(done[si], s0));
        DPR_CUDA_TRY(cudaStreamWaitEvent(st, done[si], 0));
    }
    for (int i = 1; i < NSTREAM; ++i) {
        cudaError_t e = cudaStreamSynchronize(ar->streams[i]);
        if (e != cudaSuccess && rc == DPR_OK) rc = cuda_fail(e, "cudaStreamSynchronize");
    }
    if (rc == DPR_OK) {
        // stream 0 has waited for every chunk (the fold above), so the per-pose gradients of the whole batch are complete
        if (B > 0) {
            DPR_CUDA_TRY(cudaMemcpyAsync(h_drot, g_rot, sizeof(T) * (size_t)(B * n_out * n_in), cudaMemcpyDeviceToHost, s0));
            DPR_CUDA_TRY(cudaMemcpyAsync(h_dtr, g_tr, sizeof(T) * (size_t)(B * n_out), cudaMemcpyDeviceToHost, s0));
            if (h_dbg) DPR_CUDA_TRY(cudaMemcpyAsync(h_dbg, g_bg, sizeof(T) * (size_t)B, cudaMemcpyDeviceToHost, s0));
            if (h_dow) DPR_CUDA_TRY(cudaMemcpyAsync(h_dow, g_ow, sizeof(T) * (size_t)B, cudaMemcpyDeviceToHost, s0));
        }
        if (P > 0) DPR_CUDA_TRY(cudaMemcpyAsync(h_dp, tot_dp, sizeof(T) * (size_t)(P * n_in), cudaMemcpyDeviceToHost, s0));
        if (h_dpw && P > 0) DPR_CUDA_TRY(cudaMemcpyAsync(h_dpw, tot_dpw, sizeof(T) * (size_t)P, cudaMemcpyDeviceToHost, s0));
    }
    cudaError_t e = cudaStreamSynchronize(s0);
    if (e != cudaSuccess && rc == DPR_OK) rc = cuda_fail(e, "cudaStreamSynchronize");
    return rc;
}

}  // namespace dpr

using namespace dpr;

// ---- non-blocking host-buffer calls: the blocking entry point runs on a helper thread; dpr_host_wait joins it --------
struct dpr_ticket {
    std::thread worker;
    int rc = DPR_OK;
    char message[512] = "";
};
template <typename Fn>
static int host_async(dpr_ticket_t* ticket, Fn fn) {
    if (!ticket) return DPR_ERR_NULL_POINTER;
    int dev = -1;
    DPR_CUDA_TRY(cudaGetDevice(&dev));
    dpr_ticket* t = new dpr_ticket();
    t->worker = std::thread([t, dev, fn] {
        cudaSetDevice(dev);                     // the helper thread works on the caller's device
        t->rc = fn();
        if (t->rc != DPR_OK) snprintf(t->message, sizeof(t->message), "%s", tl_error);
    });
    *ticket = t;
    return DPR_OK;
}
extern "C" {

int dpr_version(void) { return 100; }

const char* dpr_status_string(int status) {
    switch (status) {
        case DPR_OK: return "ok";
        case DPR_ERR_BAD_DIMS: return "bad dimensions (negative size, grid extent < 1, or size overflow)";
        case DPR_ERR_UNSUPPORTED: return "unsupported (N_in, N_out) or element type; supported: 1 <= N_in <= 4, 1 <= N_out <= 4 in f32/f64";
        case DPR_ERR_NULL_POINTER: return "a required pointer is NULL";
        case DPR_ERR_WORKSPACE: return "workspace smaller than dpr_workspace_bytes()";
        case DPR_ERR_CUDA: return "CUDA runtime error (see dpr_last_error_message)";
        case DPR_ERR_NO_DEVICE: return "no usable sm_100 CUDA device; libdpr has no CPU fallback";
        case DPR_ERR_NCCL: return "NCCL error";
        case DPR_ERR_BAD_OPTION: return "unknown option or bad option value";
        default: return "unknown status";
    }
}

const char* dpr_last_error_message(void) { return tl_error; }

size_t dpr_workspace_bytes(int op, int n_in, int n_out, const int64_t* grid, int64_t P, int64_t B, int sizeof_T) {
    // arguments the entry points would reject anyway: no scratch (and no division by n_in = 0 in the sort plan)
    if (n_in < 1 || n_in > kMaxDim || n_out < 1 || n_out > kMaxDim || P < 0 || B < 0 || (sizeof_T != 4 && sizeof_T != 8)) return 0;
    if (op == DPR_OP_FORWARD) return forward_workspace_bytes(n_in, n_out, grid, P, B, sizeof_T);
    if (op == DPR_OP_PULLBACK) return pullback_workspace_bytes(n_in, n_out, grid, P, B, sizeof_T);
    return 0;
}

int dpr_raster_forward_f32(int n_in, int n_out, const int64_t* grid, int64_t P, int64_t B, const float* points,
                           const float* rotation, const float* translation, const float* background,
                           const float* out_weight, const float* point_weight, float* out, void* ws, size_t ws_bytes,
                           dpr_stream_t stream) {
    return forward_entry<float>(n_in, n_out, grid, P, B, points, rotation, translation, background, out_weight, point_weight, out, ws, ws_bytes, stream);
}
int dpr_raster_forward_f64(int n_in, int n_out, const int64_t* grid, int64_t P, int64_t B, const double* points,
                           const double* rotation, const double* translation, const double* background,
                           const double* out_weight, const double* point_weight, double* out, void* ws, size_t ws_bytes,
                           dpr_stream_t stream) {
    return forward_entry<double>(n_in, n_out, grid, P, B, points, rotation, translation, background, out_weight, point_weight, out, ws, ws_bytes, stream);
}
int dpr_raster_pullback_f32(int n_in, int n_out, const int64_t* grid, int64_t P, int64_t B, const float* ds_dout,
                            const float* points, const float* rotation, const float* translation,
                            const float* out_weight, const float* point_weight, float* d_points, float* d_rotation,
                            float* d_translation, float* d_background, float* d_out_weight, float* d_point_weight,
                            void* ws, size_t ws_bytes, dpr_stream_t stream) {
    return pullback_entry<float>(n_in, n_out, grid, P, B, ds_dout, points, rotation, translation, out_weight, point_weight,
                                 d_points, d_rotation, d_translation, d_background, d_out_weight, d_point_weight, ws, ws_bytes, stream);
}
int dpr_raster_pullback_f64(int n_in, int n_out, const int64_t* grid, int64_t P, int64_t B, const double* ds_dout,
                            const double* points, const double* rotation, const double* translation,
                            const double* out_weight, const double* point_weight, double* d_points, double* d_rotation,
                            double* d_translation, double* d_background, double* d_out_weight, double* d_point_weight,
                            void* ws, size_t ws_bytes, dpr_stream_t stream) {
    return pullback_entry<double>(n_in, n_out, grid, P, B, ds_dout, points, rotation, translation, out_weight, point_weight,
                                  d_points, d_rotation, d_translation, d_background, d_out_weight, d_point_weight, ws, ws_bytes, stream);
}

int dpr_raster_forward_host_f32(int n_in, int n_out, const int64_t* grid, int64_t P, int64_t B, const float* points,
                                const float* rotation, const float* translation, const float* background,
                                const float* out_weight, const float* point_weight, float* out) {
    return forward_host<float>(n_in, n_out, grid, P, B, points, rotation, translation, background, out_weight, point_weight, out);
}
int dpr_raster_forward_host_f64(int n_in, int n_out, const int64_t* grid, int64_t P, int64_t B, const double* points,
                                const double* rotation, const double* translation, const double* background,
                                const double* out_weight, const double* point_weight, double* out) {
    return forward_host<double>(n_in, n_out, grid, P, B, points, rotation, translation, background, out_weight, point_weight, out);
}
int dpr_raster_pullback_host_f32(int n_in, int n_out, const int64_t* grid, int64_t P, int64_t B, const float* ds_dout,
                                 const float* points, const float* rotation, const float* translation,
                                 const float* out_weight, const float* point_weight, float* d_points,
                                 float* d_rotation, float* d_translation, float* d_background, float* d_out_weight,
                                 float* d_point_weight) {
    return pullback_host<float>(n_in, n_out, grid, P, B, ds_dout, points, rotation, translation, out_weight, point_weight,
                                d_points, d_rotation, d_translation, d_background, d_out_weight, d_point_weight);
}
int dpr_raster_pullback_host_f64(int n_in, int n_out, const int64_t* grid, int64_t P, int64_t B, const double* ds_dout,
                                 const double* points, const double* rotation, const double* translation,
                                 const double* out_weight, const double* point_weight, double* d_points,
                                 double* d_rotation, double* d_translation, double* d_background, double* d_out_weight,
                                 double* d_point_weight) {
    return pullback_host<double>(n_in, n_out, grid, P, B, ds_dout, points, rotation, translation, out_weight, point_weight,
                                 d_points, d_rotation, d_translation, d_background, d_out_weight, d_point_weight);
}

int dpr_host_alloc(void** ptr, size_t bytes) {
    if (!ptr) return DPR_ERR_NULL_POINTER;
    DPR_CUDA_TRY(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return DPR_OK;
}
int dpr_host_free(void* ptr) {
    if (ptr) DPR_CUDA_TRY(cudaFreeHost(ptr));
    return DPR_OK;
}
int dpr_host_release(void) {
    int dev = -1;
    DPR_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return DPR_ERR_NO_DEVICE;
    for (int op = 0; op < 2; ++op) {
        std::lock_guard<std::mutex> lock(g_arena_mutex[dev][op]);
        HostArena& a = g_arena[dev][op];
        if (a.device != dev) continue;
        for (int i = 0; i < NSTREAM; ++i) { if (a.slot[i]) cudaFree(a.slot[i]); a.slot[i] = nullptr; }
        a.slot_bytes = 0;
        for (int i = 0; i < NSTREAM; ++i) { if (a.ws[i]) cudaFree(a.ws[i]); a.ws[i] = nullptr; }
        a.ws_bytes = 0;
        for (int i = 0; i < NSTREAM; ++i) {
            if (cudaMalloc(&a.ws[i], 4096) != cudaSuccess || cudaMemset(a.ws[i], 0, 256) != cudaSuccess) return DPR_ERR_CUDA;
        }
        a.ws_bytes = 4096;
        if (a.shared) cudaFree(a.shared);
        a.shared = nullptr; a.shared_bytes = 0;
    }
    return DPR_OK;
}

int dpr_host_wait(dpr_ticket_t ticket) {
    if (!ticket) return DPR_ERR_NULL_POINTER;
    if (ticket->worker.joinable()) ticket->worker.join();
    const int rc = ticket->rc;
    if (rc != DPR_OK) set_error_message(ticket->message);
    delete ticket;
    return rc;
}
int dpr_raster_forward_host_async_f32(int n_in, int n_out, const int64_t* grid, int64_t P, int64_t B, const float* points,
                                      const float* rotation, const float* translation, const float* background,
                                      const float* out_weight, const float* point_weight, float* out, dpr_ticket_t* ticket) {
    if (!grid || n_out < 1 || n_out > kMaxDim) return DPR_ERR_BAD_DIMS;
    std::vector<int64_t> g(grid, grid + n_out);
    return host_async(ticket, [=] { return forward_host<float>(n_in, n_out, g.data(), P, B, points, rotation, translation, background, out_weight, point_weight, out); });
}
int dpr_raster_forward_host_async_f64(int n_in, int n_out, const int64_t* grid, int64_t P, int64_t B, const double* points,
                                      const double* rotation, const double* translation, const double* background,
                                      const double* out_weight, const double* point_weight, double* out, dpr_ticket_t* ticket) {
    if (!grid || n_out < 1 || n_out > kMaxDim) return DPR_ERR_BAD_DIMS;
    std::vector<int64_t> g(grid, grid + n_out);
    return host_async(ticket, [=] { return forward_host<double>(n_in, n_out, g.data(), P, B, points, rotation, translation, background, out_weight, point_weight, out); });
}
int dpr_raster_pullback_host_async_f32(int n_in, int n_out, const int64_t* grid, int64_t P, int64_t B, const float* ds_dout,
                                       const float* points, const float* rotation, const float* translation,
                                       const float* out_weight, const float* point_weight, float* d_points,
                                       float* d_rotation, float* d_translation, float* d_background, float* d_out_weight,
                                       float* d_point_weight, dpr_ticket_t* ticket) {
    if (!grid || n_out < 1 || n_out > kMaxDim) return DPR_ERR_BAD_DIMS;
    std::vector<int64_t> g(grid, grid + n_out);
    return host_async(ticket, [=] { return pullback_host<float>(n_in, n_out, g.data(), P, B, ds_dout, points, rotation, translation, out_weight, point_weight,
                                                                d_points, d_rotation, d_translation, d_background, d_out_weight, d_point_weight); });
}
int dpr_raster_pullback_host_async_f64(int n_in, int n_out, const int64_t* grid, int64_t P, int64_t B, const double* ds_dout,
                                       const double* points, const double* rotation, const double* translation,
                                       const double* out_weight, const double* point_weight, double* d_points,
                                       double* d_rotation, double* d_translation, double* d_background, double* d_out_weight,
                                       double* d_point_weight, dpr_ticket_t* ticket) {
    if (!grid || n_out < 1 || n_out > kMaxDim) return DPR_ERR_BAD_DIMS;
    std::vector<int64_t> g(grid, grid + n_out);
    return host_async(ticket, [=] { return pullback_host<double>(n_in, n_out, g.data(), P, B, ds_dout, points, rotation, translation, out_weight, point_weight,
                                                                 d_points, d_rotation, d_translation, d_background, d_out_weight, d_point_weight); });
}

int dpr_set_option(int option, int64_t value) {
    if (value < 0) return DPR_ERR_BAD_OPTION;
    switch (option) {
        case DPR_OPT_FORWARD_ALGO: if (value > 3) return DPR_ERR_BAD_OPTION; g_tuning.forward_algo = value; return DPR_OK;
        case DPR_OPT_PULLBACK_ALGO: if (value > 7 || value == 5 || value == 6) return DPR_ERR_BAD_OPTION; g_tuning.pullback_algo = value; return DPR_OK;
        case DPR_OPT_TILE_SMEM_BYTES: g_tuning.tile_smem_bytes = value; return DPR_OK;
        case DPR_OPT_POINT_SPLIT: g_tuning.point_split = value; return DPR_OK;
        case DPR_OPT_POSE_CHUNK: g_tuning.pose_chunk = value; return DPR_OK;
        case DPR_OPT_FORWARD_ACCUM: if (value > 1) return DPR_ERR_BAD_OPTION; g_tuning.forward_accum = value; return DPR_OK;
        case DPR_OPT_POINT_SORT: if (value > 2) return DPR_ERR_BAD_OPTION; g_tuning.point_sort = value; return DPR_OK;
        case DPR_OPT_TILE3D_TMA: if (value > 1) return DPR_ERR_BAD_OPTION; g_tuning.tile3d_tma = value; return DPR_OK;
        case DPR_OPT_BINNING_CACHE: if (value > 1) return DPR_ERR_BAD_OPTION; g_tuning.binning_cache = value; return DPR_OK;
        case DPR_OPT_COMM_P2P: if (value > 1) return DPR_ERR_BAD_OPTION; g_tuning.comm_p2p = value; return DPR_OK;
        default: return DPR_ERR_BAD_OPTION;
    }
}
int64_t dpr_get_option(int option) {
    switch (option) {
        case DPR_OPT_FORWARD_ALGO: return g_tuning.forward_algo;
        case DPR_OPT_PULLBACK_ALGO: return g_tuning.pullback_algo;
        case DPR_OPT_TILE_SMEM_BYTES: return g_tuning.tile_smem_bytes;
        case DPR_OPT_POINT_SPLIT: return g_tuning.point_split;
        case DPR_OPT_POSE_CHUNK: return g_tuning.pose_chunk;
        case DPR_OPT_FORWARD_ACCUM: return g_tuning.forward_accum;
        case DPR_OPT_POINT_SORT: return g_tuning.point_sort;
        case DPR_OPT_TILE3D_TMA: return g_tuning.tile3d_tma;
        case DPR_OPT_BINNING_CACHE: return g_tuning.binning_cache;
        case DPR_OPT_COMM_P2P: return g_tuning.comm_p2p;
        default: return -1;
    }
}
int dpr_profile_enable(int on) {
    profile_clear();
    g_prof_on.store(on == 2 ? 2 : (on ? 1 : 0));
    return DPR_OK;
}
int dpr_profile_count(void) {
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    return (int)g_prof.size();
}
int dpr_profile_get(int i, const char** name, float* ms) {
    ProfileRecord r;
    {
        std::lock_guard<std::mutex> lock(g_prof_mutex);
        if (i < 0 || i >= (int)g_prof.size()) return DPR_ERR_BAD_OPTION;
        r = g_prof[i];
    }
    DPR_CUDA_TRY(cudaEventSynchronize(r.e1));
    float t = 0.f;
    DPR_CUDA_TRY(cudaEventElapsedTime(&t, r.e0, r.e1));
    if (name) *name = r.name;
    if (ms) *ms = t;
    return DPR_OK;
}
int64_t dpr_kernel_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
const char* dpr_last_path(int op) { return (op >= 0 && op < 2) ? tl_path[op] : "none"; }

}  // extern "C"
