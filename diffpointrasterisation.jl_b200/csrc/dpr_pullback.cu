// dpr_pullback.cu - batched pullback of the splat for sm_100a.
//
// Replaces the reference's CUDA pullback (ext/DiffPointRasterisationCUDAExt.jl:19-210 kernel, :231-321 driver)
// and computes the same gradients as the CPU method (src/raster_pullback.jl:2-82 per pose, :85-148 batched):
// gather ds_dout at the 2^N_out stencil cells of every (point, pose) pair and reduce to
//   d_points (summed over poses), d_point_weight (summed over poses),
//   d_rotation, d_translation, d_out_weight (summed over points, per pose), d_background (sum of ds_dout, per pose).
//
// Decomposition: a thread OWNS K points and loops over the poses of its CTA's pose chunk, so
// the pose-sum of d_points / d_point_weight is a register accumulation (one REDG per point and pose chunk, not per
// splat); the per-pose point-sums are reduced with warp shuffles, combined per CTA in shared memory, and leave the
// CTA as one REDG per (pose, value).  CTAs are ordered so that neighbours work on the same poses at the same time,
// which keeps the ds_dout images they gather from resident in L1/L2.
#include <cmath>
#include <cstring>

#include "dpr_common.cuh"
#include "dpr_internal.h"
#include "dpr_sort.cuh"
#include "dpr_tile3d.cuh"

namespace dpr {

template <int N_IN, int N_OUT>
struct PoseGradLayout {
    // per-pose values reduced over points: d_rotation (N_OUT*N_IN, column-major), d_translation (N_OUT), d_out_weight
    static constexpr int NV = N_OUT * N_IN + N_OUT + 1;
};

// Gradient pieces of one (point, pose) pair.  `img` is the pose's ds_dout image.
//   s      = sum_c W_c G_c                                 (src/raster_pullback.jl:55-58)
//   gk[n]  = sum_c G_c * sign_n(c) * prod_{m != n} w_m(c)  (src/raster_pullback.jl:60-65, :150-160)
// Out-of-bounds corners are skipped individually (src/raster_pullback.jl:51).
template <typename T, int N_OUT, typename Fetch>
__device__ __forceinline__ void gather_corners(const Grid<T, N_OUT>& grid, const int (&i0)[N_OUT], const T (&dl)[N_OUT],
                                               Fetch fetch, T& s, T (&gk)[N_OUT]) {
    T du[N_OUT];
#pragma unroll
    for (int k = 0; k < N_OUT; ++k) { du[k] = T(1) - dl[k]; gk[k] = T(0); }
    s = T(0);
#pragma unroll
    for (int c = 0; c < (1 << N_OUT); ++c) {
        bool inb = true;
        int64_t off = 0, stride = 1;
#pragma unroll
        for (int k = 0; k < N_OUT; ++k) {
            const int idx = i0[k] + ((c >> k) & 1);
            inb = inb && idx >= 0 && idx < grid.g[k];
            off += idx * stride;
            stride *= grid.g[k];
        }
        const T G = inb ? fetch(off) : T(0);
        s += corner_weight<T, N_OUT>(c, dl, du) * G;
#pragma unroll
        for (int n = 0; n < N_OUT; ++n) {
            T iw = ((c >> n) & 1) ? T(1) : T(-1);
#pragma unroll
            for (int m = 0; m < N_OUT; ++m)
                if (m != n) iw *= ((c >> m) & 1) ? dl[m] : du[m];
            gk[n] += G * iw;
        }
    }
}

template <typename T, int N_IN, int N_OUT, int K>
__global__ void __launch_bounds__(256)
pullback_gather_global_kernel(const T* __restrict__ ds_dout, const T* __restrict__ points,
                              const T* __restrict__ rotation, const T* __restrict__ translation,
                              const T* __restrict__ out_weight, const T* __restrict__ point_weight,
                              T* __restrict__ d_points, T* __restrict__ d_rotation, T* __restrict__ d_translation,
                              T* __restrict__ d_out_weight, T* __restrict__ d_point_weight,
                              const int32_t* __restrict__ perm, Grid<T, N_OUT> grid, int64_t P, int64_t B,
                              int point_chunks, int pose_chunk) {
    constexpr int NV = PoseGradLayout<N_IN, N_OUT>::NV;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* pose_acc = reinterpret_cast<T*>(smem_raw);  // [pose_chunk][NV]

    const int pc = blockIdx.x % point_chunks;
    const int64_t bc = blockIdx.x / point_chunks;
    const int64_t b0 = bc * pose_chunk;
    const int64_t b1 = (b0 + pose_chunk < B) ? b0 + pose_chunk : B;
    const int n_pose = (int)(b1 - b0);
    for (int i = threadIdx.x; i < n_pose * NV; i += blockDim.x) pose_acc[i] = T(0);
    __syncthreads();

    T x[K][N_IN], pw[K], dpt[K][N_IN], dpw[K];
    bool valid[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int64_t p = ((int64_t)pc * K + k) * blockDim.x + threadIdx.x;
        valid[k] = p < P;
        const int64_t pp = valid[k] ? p : 0;
        load_point(x[k], points, pp);
        pw[k] = point_weight ? __ldg(point_weight + pp) : T(1);
#pragma unroll
        for (int j = 0; j < N_IN; ++j) dpt[k][j] = T(0);
        dpw[k] = T(0);
    }

    const int lane = threadIdx.x & 31;
    for (int64_t b = b0; b < b1; ++b) {
        Pose<T, N_IN, N_OUT> pose;
        load_pose(pose, rotation, translation, out_weight, b);
        const T* __restrict__ img = ds_dout + b * grid.cells;
        T acc[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] = T(0);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            int i0[N_OUT];
            T dl[N_OUT];
            if (!valid[k] || !stencil(x[k], pose, grid, i0, dl)) continue;
            T s, gk[N_OUT];
            gather_corners<T, N_OUT>(grid, i0, dl, [&](int64_t off) { return __ldg(img + off); }, s, gk);
            acc[NV - 1] += s * pw[k];                         // d_out_weight,   src/raster_pullback.jl:57
            dpw[k] += s * pose.ow;                            // d_point_weight, src/raster_pullback.jl:58
            const T f = pose.ow * pw[k];                      // factor / G,     src/raster_pullback.jl:60
            T scaled[N_OUT];
#pragma unroll
            for (int n = 0; n < N_OUT; ++n) {
                scaled[n] = (f * gk[n]) * grid.scale[n];      // src/raster_pullback.jl:67
                acc[N_OUT * N_IN + n] += scaled[n];           // d_translation,  src/raster_pullback.jl:68
            }
#pragma unroll
            for (int j = 0; j < N_IN; ++j) {
                T d = T(0);
#pragma unroll
                for (int n = 0; n < N_OUT; ++n) {
                    acc[n + j * N_OUT] += scaled[n] * x[k][j];   // d_rotation, src/raster_pullback.jl:69
                    d += pose.R[n][j] * scaled[n];               // R' * scaled, src/raster_pullback.jl:70
                }
                dpt[k][j] += d;                                   // src/raster_pullback.jl:71
            }
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const T r = warp_sum(acc[v]);
            if (lane == 0 && r != T(0)) atomicAdd(&pose_acc[(int)(b - b0) * NV + v], r);
        }
    }

#pragma unroll
    for (int k = 0; k < K; ++k) {
        if (!valid[k]) continue;
        int64_t p = ((int64_t)pc * K + k) * blockDim.x + threadIdx.x;
        if (perm) p = __ldg(perm + p);
#pragma unroll
        for (int j = 0; j < N_IN; ++j) red_add(d_points + p * N_IN + j, dpt[k][j]);
        if (d_point_weight) red_add(d_point_weight + p, dpw[k]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_pose * NV; i += blockDim.x) {
        const int bl = i / NV, v = i % NV;
        const T r = pose_acc[i];
        const int64_t b = b0 + bl;
        if (v < N_OUT * N_IN) red_add(d_rotation + b * (N_OUT * N_IN) + v, r);
        else if (v < N_OUT * N_IN + N_OUT) red_add(d_translation + b * N_OUT + (v - N_OUT * N_IN), r);
        else if (d_out_weight) red_add(d_out_weight + b, r);
    }
}

// d_background[b] = sum(ds_dout[:, b])   (src/raster_pullback.jl:78; ext/DiffPointRasterisationCUDAExt.jl:265-267)
template <typename T>
__global__ void __launch_bounds__(256)
background_sum_kernel(const T* __restrict__ ds_dout, T* __restrict__ d_background, int64_t cells, int segs,
                      int64_t seg_len) {
    constexpr int VEC = 16 / sizeof(T);
    struct alignas(16) Pack { T v[VEC]; };
    const int64_t b = blockIdx.x / segs;
    const int seg = blockIdx.x % segs;
    const int64_t lo = (int64_t)seg * seg_len;
    const int64_t hi = (lo + seg_len < cells) ? lo + seg_len : cells;
    const T* __restrict__ src = ds_dout + b * cells;
    T s = T(0);
    if ((cells % VEC) == 0 && (seg_len % VEC) == 0 && (reinterpret_cast<uintptr_t>(ds_dout) % 16) == 0) {
        const Pack* __restrict__ v = reinterpret_cast<const Pack*>(src);
        for (int64_t i = lo / VEC + threadIdx.x; i < hi / VEC; i += blockDim.x) {
            const Pack pk = v[i];
#pragma unroll
            for (int k = 0; k < VEC; ++k) s += pk.v[k];
        }
    } else {
        for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) s += src[i];
    }
    __shared__ T warp_part[8];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        T t = threadIdx.x < 8 ? warp_part[threadIdx.x] : T(0);
        t = warp_sum(t);
        if (threadIdx.x == 0) {
            if (segs == 1) d_background[b] = t;
            else red_add(d_background + b, t);
        }
    }
}

template <typename T>
static int launch_background_sum(const PullbackArgs<T>& a, int64_t cells, const DeviceInfo& dev) {
    if (!a.d_background || a.B == 0) return DPR_OK;
    int64_t segs = 1;
    const int64_t min_seg = 256 * 16 * 4;  // keep >= 64 KB (f32) of work per CTA
    if (a.B < (int64_t)dev.sm_count * 4) {
        segs = ((int64_t)dev.sm_count * 4 + a.B - 1) / a.B;
        const int64_t max_segs = (cells + min_seg - 1) / min_seg;
        if (segs > max_segs) segs = max_segs;
        if (segs < 1) segs = 1;
    }
    int64_t seg_len = (cells + segs - 1) / segs;
    seg_len = (seg_len + 15) / 16 * 16;
    segs = (cells + seg_len - 1) / seg_len;
    if (segs < 1) segs = 1;
    if (segs > 1) DPR_CUDA_TRY(cudaMemsetAsync(a.d_background, 0, sizeof(T) * (size_t)a.B, a.stream));
    {
        LaunchScope scope("background_sum", a.stream);
        background_sum_kernel<T><<<(unsigned)(a.B * segs), 256, 0, a.stream>>>(a.ds_dout, a.d_background, cells, (int)segs, seg_len);
    }
    DPR_CUDA_TRY(cudaGetLastError());
    return DPR_OK;
}

// Spatial sort of the points pays off once the sort (three small kernels) is amortised over enough poses and there
// are enough points for a warp's 32 points to be neighbours; it needs the workspace of dpr_workspace_bytes().
template <typename T>
static bool use_sort(const PullbackArgs<T>& a) {
    if (tuning().point_sort == 2) return false;
    if (a.P >= (int64_t)0x7fffffff) return false;
    const SortPlan sp = make_sort_plan(a.n_in, a.P, (int)sizeof(T), a.point_weight != nullptr, 256);
    if (!a.workspace || a.workspace_bytes < sp.total) return false;
    if (tuning().point_sort == 1) return true;
    return a.P >= 4096 && a.B >= 4 && a.P * a.B >= ((int64_t)1 << 22);   // three extra launches must pay for themselves
}

template <typename T, int N_IN, int N_OUT>
static int pullback_global(const PullbackArgs<T>& a, const DeviceInfo& dev) {
    constexpr int NV = PoseGradLayout<N_IN, N_OUT>::NV;
    // points owned per thread: 4 for the tuned dimensions, fewer when N_in * N_out values per point would spill
    constexpr int K = (N_IN <= 3 && N_OUT <= 3) ? 4 : ((N_IN == 4 && N_OUT == 4) ? 1 : 2);
    Grid<T, N_OUT> grid;
    grid.cells = 1;
    for (int k = 0; k < N_OUT; ++k) {
        grid.g[k] = (int)a.grid[k];
        grid.scale[k] = T(a.grid[k]) / T(2);  // src/raster_pullback.jl:29
        grid.cells *= a.grid[k];
    }
    {
        int rcz = zero_gradients(a);
        if (rcz != DPR_OK) return rcz;
    }
    int rc = launch_background_sum(a, grid.cells, dev);
    if (rc != DPR_OK) return rc;
    if (a.P == 0 || a.B == 0) return DPR_OK;
    const T* pts = a.points;
    const T* pwt = a.point_weight;
    const int32_t* perm = nullptr;
    if (use_sort(a)) {
        const SortPlan sp = make_sort_plan(N_IN, a.P, (int)sizeof(T), a.point_weight != nullptr, 256);
        rc = sort_points<T, N_IN>(a.points, a.point_weight, a.P, a.workspace, sp, dev, a.stream);
        if (rc != DPR_OK) return rc;
        char* ws = static_cast<char*>(a.workspace);
        pts = reinterpret_cast<const T*>(ws + sp.off_points);
        if (a.point_weight) pwt = reinterpret_cast<const T*>(ws + sp.off_pw);
        perm = reinterpret_cast<const int32_t*>(ws + sp.off_perm);
    }

    const int threads = 256;
    const int64_t point_chunks = (a.P + (int64_t)threads * K - 1) / ((int64_t)threads * K);
    // poses per CTA: enough CTAs for ~4 waves of 8 CTAs/SM, but long enough loops to amortise the d_points REDGs
    int64_t pose_chunk = tuning().pose_chunk;
    if (pose_chunk <= 0) {
        const int64_t want_ctas = (int64_t)dev.sm_count * 8 * 4;
        int64_t pose_chunks = (want_ctas + point_chunks - 1) / point_chunks;
        if (pose_chunks < 1) pose_chunks = 1;
        if (pose_chunks > a.B) pose_chunks = a.B;
        pose_chunk = (a.B + pose_chunks - 1) / pose_chunks;
        if (pose_chunk < 8) pose_chunk = a.B < 8 ? a.B : 8;
    }
    if (pose_chunk > 512) pose_chunk = 512;
    // the per-pose accumulators must fit the 48 KB of dynamic shared memory a kernel gets without opting in
    // (Float64 with N_in * N_out >= 9: 512 poses would need 53 - 86 KB and the launch would fail)
    if (pose_chunk > (int64_t)(48 * 1024 / (sizeof(T) * NV))) pose_chunk = (int64_t)(48 * 1024 / (sizeof(T) * NV));
    if (pose_chunk > a.B) pose_chunk = a.B;
    const int64_t pose_chunks = (a.B + pose_chunk - 1) / pose_chunk;
    if (point_chunks * pose_chunks > (int64_t)0x7fffffff || point_chunks > (int64_t)0x7fffffff) return DPR_ERR_BAD_DIMS;
    const size_t smem = sizeof(T) * (size_t)pose_chunk * NV;
    {
        LaunchScope scope("pullback_gather_global", a.stream);
        pullback_gather_global_kernel<T, N_IN, N_OUT, K><<<(unsigned)(point_chunks * pose_chunks), threads, smem, a.stream>>>(
            a.ds_dout, pts, a.rotation, a.translation, a.out_weight, pwt, a.d_points, a.d_rotation,
            a.d_translation, a.d_out_weight, a.d_point_weight, perm, grid, a.P, a.B, (int)point_chunks, (int)pose_chunk);
    }
    DPR_CUDA_TRY(cudaGetLastError());
    set_last_path(DPR_OP_PULLBACK, perm ? "gather_global_sorted" : "gather_global");
    return DPR_OK;
}

}  // namespace dpr
#include "dpr_pullback_fast.cuh"
#include "dpr_pullback_tma.cuh"
namespace dpr {

// zero everything that is accumulated with REDG (the five fill! calls of ext/DiffPointRasterisationCUDAExt.jl:272-276)
// in ONE launch: for small problems the step is launch bound, five memsets are five launches
struct ZeroList {
    void* ptr[5];
    unsigned long long bytes[5];   // each a multiple of 4
};
__global__ void __launch_bounds__(256) zero_buffers_kernel(ZeroList z) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        uint32_t* p = static_cast<uint32_t*>(z.ptr[k]);
        const size_t n = (size_t)(z.bytes[k] / 4);
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = 0u;
    }
}

template <typename T>
static int zero_gradients(const PullbackArgs<T>& a) {
    const int nr = a.n_in * a.n_out;
    ZeroList z;
    z.ptr[0] = a.d_points;       z.bytes[0] = sizeof(T) * (unsigned long long)(a.P * a.n_in);
    z.ptr[1] = a.d_rotation;     z.bytes[1] = sizeof(T) * (unsigned long long)(a.B * nr);
    z.ptr[2] = a.d_translation;  z.bytes[2] = sizeof(T) * (unsigned long long)(a.B * a.n_out);
    z.ptr[3] = a.d_out_weight;   z.bytes[3] = a.d_out_weight ? sizeof(T) * (unsigned long long)a.B : 0;
    z.ptr[4] = a.d_point_weight; z.bytes[4] = a.d_point_weight ? sizeof(T) * (unsigned long long)a.P : 0;
    unsigned long long most = 0;
    for (int k = 0; k < 5; ++k) most = z.bytes[k] > most ? z.bytes[k] : most;
    if (most == 0) return DPR_OK;
    unsigned long long blocks = (most / 4 + 255) / 256;
    if (blocks > 592) blocks = 592;
    if (blocks < 1) blocks = 1;
    {
        LaunchScope scope("zero_gradients", a.stream);
        zero_buffers_kernel<<<(unsigned)blocks, 256, 0, a.stream>>>(z);
    }
    DPR_CUDA_TRY(cudaGetLastError());
    return DPR_OK;
}

// ---------------------------------------------------------------------------------------------------------
// 3-d grids: per-pose tile binning + one CTA per (pose, tile) gathering from its staged ds_dout tile (dpr_tile3d.cuh)
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tensor_map_encoder() {
    static const EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        cudaGetLastError();
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

template <typename T, int N_IN>
static int pullback_tile3d(const PullbackArgs<T>& a, const DeviceInfo& dev, const t3::Plan& pl) {
    constexpr int NR = 3 * N_IN;
    const Grid<T, 3> grid = t3::make_grid3<T>(a.grid);
    char* ws = static_cast<char*>(a.workspace);
    {   // per-pose outputs are accumulated with REDG; d_points / d_point_weight are overwritten by the un-permute pass
        ZeroList z;
        z.ptr[0] = a.d_rotation;    z.bytes[0] = sizeof(T) * (unsigned long long)(a.B * NR);
        z.ptr[1] = a.d_translation; z.bytes[1] = sizeof(T) * (unsigned long long)(a.B * 3);
        z.ptr[2] = a.d_out_weight;  z.bytes[2] = a.d_out_weight ? sizeof(T) * (unsigned long long)a.B : 0;
        z.ptr[3] = a.d_background;  z.bytes[3] = a.d_background ? sizeof(T) * (unsigned long long)a.B : 0;
        z.ptr[4] = nullptr;         z.bytes[4] = 0;
        LaunchScope scope("zero_gradients", a.stream);
        zero_buffers_kernel<<<8, 256, 0, a.stream>>>(z);
    }
    t3::CacheCtl cache;
    int rc = t3::cache_begin<T>(cache, ws, pl, N_IN, a.grid, a.points, a.point_weight, a.rotation, a.translation, a.P, a.B, dev, a.stream);
    if (rc != DPR_OK) return rc;
    rc = t3::presort<T, N_IN>(a.points, a.point_weight, a.P, ws, pl, cache, dev, a.stream);
    if (rc != DPR_OK) return rc;
    DPR_CUDA_TRY(cudaMemsetAsync(ws + pl.off_acc4, 0, sizeof(T) * 4 * (size_t)a.P, a.stream));     // packed pose-sum accumulator
    // tensor-map TMA for the tile loads when the volume qualifies (16-byte rows and base): one 4-d map over
    // (g0, g1, g2, B), box = one tile.  Tile origins are multiples of 32 cells, which satisfies the 16-byte box-start rule
    // (dpr_tile3d.cuh); DPR_OPT_TILE3D_TMA = 1 forces the cooperative loads.
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    bool use_tma = false;
    if (tuning().tile3d_tma != 1 && (a.grid[0] * sizeof(T)) % 16 == 0 && (reinterpret_cast<uintptr_t>(a.ds_dout) % 16) == 0 &&
        a.B < ((int64_t)1 << 31) && tensor_map_encoder()) {
        const cuuint64_t dims[4] = {(cuuint64_t)a.grid[0], (cuuint64_t)a.grid[1], (cuuint64_t)a.grid[2], (cuuint64_t)a.B};
        const cuuint64_t strides[3] = {(cuuint64_t)a.grid[0] * sizeof(T), (cuuint64_t)a.grid[0] * a.grid[1] * sizeof(T),
                                       (cuuint64_t)grid.cells * sizeof(T)};
        const cuuint32_t box[4] = {(cuuint32_t)t3::TX, (cuuint32_t)t3::TY, (cuuint32_t)t3::TZ, 1u};
        const cuuint32_t es[4] = {1u, 1u, 1u, 1u};
        const CUresult r = tensor_map_encoder()(&map, sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4,
                                                const_cast<T*>(a.ds_dout), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        use_tma = r == CUDA_SUCCESS;
    }
    const size_t tile_bytes = sizeof(T) * (size_t)t3::kTileCells;
    const t3::Pt4<T>* pts4 = reinterpret_cast<const t3::Pt4<T>*>(ws + pl.off_pts4);
    t3::Pt4<T>* acc4 = reinterpret_cast<t3::Pt4<T>*>(ws + pl.off_acc4);
    const uint32_t* cnt = reinterpret_cast<const uint32_t*>(ws + pl.tile_scan.off_data);
    const uint32_t* entries = reinterpret_cast<const uint32_t*>(ws + pl.off_entries);
    auto launch = [&](auto kern, int64_t b0, int64_t nb) -> int {
        const int rcs = opt_in_smem_once(kern, tile_bytes, dev);
        if (rcs != DPR_OK) return rcs;
        LaunchScope scope("pullback_tile3d", a.stream);
        kern<<<dim3((unsigned)pl.tg.nt[0], (unsigned)pl.tg.nt[1], (unsigned)(pl.tg.nt[2] * nb)), t3::kThreads, tile_bytes, a.stream>>>(map, a.ds_dout, pts4, entries, cnt, a.rotation,
                                                                                    a.translation, a.out_weight, acc4, a.d_rotation,
                                                                                    a.d_translation, a.d_background, a.d_out_weight, grid,
                                                                                    pl.tg, b0, cache.valid);
        return DPR_OK;
    };
    for (int64_t b0 = 0; b0 < a.B; b0 += pl.group) {
        const int64_t nb = (b0 + pl.group < a.B) ? pl.group : a.B - b0;
        rc = t3::bin_poses<T, N_IN>(a.rotation, a.translation, grid, a.P, b0, nb, ws, pl, cache, a.stream);
        if (rc != DPR_OK) return rc;
        rc = use_tma ? launch(t3::pullback_tile3d_kernel<T, N_IN, true>, b0, nb) : launch(t3::pullback_tile3d_kernel<T, N_IN, false>, b0, nb);
        if (rc != DPR_OK) return rc;
    }
    {
        int64_t blocks = (a.P + 255) / 256;
        if (blocks > (int64_t)dev.sm_count * 16) blocks = (int64_t)dev.sm_count * 16;
        LaunchScope scope("tile3_unpermute", a.stream);
        t3::unpermute_kernel<T, N_IN><<<(unsigned)blocks, 256, 0, a.stream>>>(acc4, reinterpret_cast<const int32_t*>(ws + pl.off_perm), a.P,
                                                                            a.d_points, a.d_point_weight);
    }
    DPR_CUDA_TRY(cudaGetLastError());
    set_last_path(DPR_OP_PULLBACK, use_tma ? "tile3d_binned_tma" : "tile3d_binned_coop");
    return DPR_OK;
}

template <typename T, int N_IN>
static int pullback_gather2d(const PullbackArgs<T>& a, const DeviceInfo& dev) {
    constexpr int K = 4;
    constexpr int NV = 2 * N_IN + 3, PP = (NV + 3) / 4 * 4;
    Grid<T, 2> grid;
    grid.cells = 1;
    for (int k = 0; k < 2; ++k) {
        grid.g[k] = (int)a.grid[k];
        grid.scale[k] = T(a.grid[k]) / T(2);  // src/raster_pullback.jl:29
        grid.cells *= a.grid[k];
    }
    int rc = zero_gradients(a);
    if (rc != DPR_OK) return rc;
    if (a.P == 0 || a.B == 0) return launch_background_sum(a, grid.cells, dev);
    const T* pts = a.points;
    const T* pwt = a.point_weight;
    const int32_t* perm = nullptr;
    if (use_sort(a)) {
        const SortPlan sp = make_sort_plan(N_IN, a.P, (int)sizeof(T), a.point_weight != nullptr, 256);
        rc = sort_points<T, N_IN>(a.points, a.point_weight, a.P, a.workspace, sp, dev, a.stream);
        if (rc != DPR_OK) return rc;
        char* ws = static_cast<char*>(a.workspace);
        pts = reinterpret_cast<const T*>(ws + sp.off_points);
        if (a.point_weight) pwt = reinterpret_cast<const T*>(ws + sp.off_pw);
        perm = reinterpret_cast<const int32_t*>(ws + sp.off_perm);
    }
    const int threads = 256;
    const int64_t point_chunks = (a.P + (int64_t)threads * K - 1) / ((int64_t)threads * K);
    int64_t pose_chunk = tuning().pose_chunk;
    if (pose_chunk <= 0) {
        const int64_t want_ctas = (int64_t)dev.sm_count * 8 * 4;
        int64_t pose_chunks = (want_ctas + point_chunks - 1) / point_chunks;
        if (pose_chunks < 1) pose_chunks = 1;
        if (pose_chunks > a.B) pose_chunks = a.B;
        pose_chunk = (a.B + pose_chunks - 1) / pose_chunks;
        if (pose_chunk < 8) pose_chunk = a.B < 8 ? a.B : 8;
    }
    if (pose_chunk > 512) pose_chunk = 512;
    // pose records + accumulators must fit the 48 KB of dynamic shared memory a kernel gets without opting in
    // (Float64, 3-d points: 512 poses would need 86 KB and the launch would fail)
    if (pose_chunk > (int64_t)(48 * 1024 / (sizeof(T) * (PP + NV)))) pose_chunk = (int64_t)(48 * 1024 / (sizeof(T) * (PP + NV)));
    if (pose_chunk > a.B) pose_chunk = a.B;
    const int64_t pose_chunks = (a.B + pose_chunk - 1) / pose_chunk;
    // d_background: a few extra CTAs per pose chunk inside the gather launch (see the kernel) when there is enough
    // gather work to hide them behind, else the separate pass
    // (one bg CTA sums whole images, about 3 MB of them - well below the duration of a gather CTA; images above 8 MB, where
    // one image alone is a long CTA, and launches with few gather CTAs keep the separate pass: config 4 lost 1.7 ms with
    // 51 MB per bg CTA)
    int bg_ctas = 0;
    const int64_t img_bytes = grid.cells * (int64_t)sizeof(T);
    if (a.d_background && img_bytes <= ((int64_t)8 << 20) && point_chunks >= 8 && pose_chunk >= 8) {
        int64_t n = (pose_chunk * img_bytes + ((int64_t)3 << 20) - 1) / ((int64_t)3 << 20);
        if (n < 1) n = 1;
        if (n > pose_chunk) n = pose_chunk;
        if (n * 4 <= point_chunks) bg_ctas = (int)n;
    }
    if (!bg_ctas) {
        rc = launch_background_sum(a, grid.cells, dev);
        if (rc != DPR_OK) return rc;
    }
    if ((point_chunks + bg_ctas) * pose_chunks > (int64_t)0x7fffffff) return DPR_ERR_BAD_DIMS;
    size_t smem = sizeof(T) * (size_t)pose_chunk * (PP + NV);
    if (smem < sizeof(T) * (size_t)(pose_chunk * PP + 8)) smem = sizeof(T) * (size_t)(pose_chunk * PP + 8);
    auto launch = [&](auto kern) -> int {
        LaunchScope scope("pullback_gather2d", a.stream);
        kern<<<(unsigned)((point_chunks + bg_ctas) * pose_chunks), threads, smem, a.stream>>>(
            a.ds_dout, pts, a.rotation, a.translation, a.out_weight, pwt, a.d_points, a.d_rotation,
            a.d_translation, a.d_out_weight, a.d_point_weight, perm, grid, (int)a.P, a.B, (int)point_chunks, (int)pose_chunk,
            a.d_background, bg_ctas);
        return DPR_OK;
    };
    // Paired 8-byte loads (8-byte aligned rows) pay off only for UNSORTED points, where the 32 lanes of a gather hit 32
    // unrelated lines (config 2 unsorted: 2.94 ms against 3.24 ms with scalar loads).  With spatially sorted points the
    // lanes share lines and four 4-byte loads are cheaper for the L1 data stage than an 8-byte load per row plus a
    // 4-byte load on the odd lanes: config 2 2.11 -> 1.90 ms, config 4 4.20 -> 3.33 ms (profiles/probe_layout_r01.json
    // shows the same ordering in isolation).  pullback_algo 2 / 3 force pairs / scalars.
    const int64_t palgo = tuning().pullback_algo;
    const bool pair = sizeof(T) == 4 && (a.grid[0] % 2) == 0 && (reinterpret_cast<uintptr_t>(a.ds_dout) % 8) == 0 &&
                      palgo != 3 && (palgo == 2 || perm == nullptr);
    if (pair) rc = a.point_weight ? launch(pullback_gather2d_kernel<T, N_IN, K, true, true>) : launch(pullback_gather2d_kernel<T, N_IN, K, false, true>);
    else rc = a.point_weight ? launch(pullback_gather2d_kernel<T, N_IN, K, true, false>) : launch(pullback_gather2d_kernel<T, N_IN, K, false, false>);
    if (rc != DPR_OK) return rc;
    DPR_CUDA_TRY(cudaGetLastError());
    set_last_path(DPR_OP_PULLBACK, perm ? (pair ? "gather2d_pair_sorted" : "gather2d_sorted") : (pair ? "gather2d_pair" : "gather2d"));
    return DPR_OK;
}

// Float32 images that fit a 2- or 3-stage shared-memory ring: TMA-staged kernel (dpr_pullback_tma.cuh).
// `padded`: rows are fetched through a tensor map whose box is 4 columns wider than the image (bank-skewed pitch).
// padded: the staged image sits in a frame of zeros - 4 columns in front of every row, rows -2, -1, g1, g1 + 1, 4 guard words
// (dpr_pullback_tma.cuh)
static bool tma2d_can_pad(const int64_t* grid) { return (grid[0] % 4) == 0 && grid[0] + 4 <= 256 && grid[1] + 4 <= 256; }
static int64_t tma2d_stage_words(const int64_t* grid, bool padded) { return padded ? (grid[0] + 4) * (grid[1] + 4) + 4 : grid[0] * grid[1]; }

#ifndef DPR_TMA2D_K
#define DPR_TMA2D_K 8          /* points a consumer thread owns (config 5: 8 -> 5.55 ms, 10 -> 5.65 ms, 12 spills) */
#endif
template <int N_IN>
static int pullback_tma2d(const PullbackArgs<float>& a, const DeviceInfo& dev, int stages, bool padded) {
    constexpr int K = DPR_TMA2D_K;
    Grid<float, 2> grid;
    grid.cells = 1;
    for (int k = 0; k < 2; ++k) {
        grid.g[k] = (int)a.grid[k];
        grid.scale[k] = float(a.grid[k]) / 2.f;
        grid.cells *= a.grid[k];
    }
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    if (padded) {
        const cuuint64_t dims[3] = {(cuuint64_t)a.grid[0], (cuuint64_t)a.grid[1], (cuuint64_t)a.B};
        const cuuint64_t strides[2] = {(cuuint64_t)a.grid[0] * 4, (cuuint64_t)grid.cells * 4};
        const cuuint32_t box[3] = {(cuuint32_t)(a.grid[0] + 4), (cuuint32_t)(a.grid[1] + 4), 1u};      // starts at (-4, -2)
        const cuuint32_t es[3] = {1u, 1u, 1u};
        EncodeTiledFn enc = tensor_map_encoder();
        if (!enc || enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(a.ds_dout), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return DPR_ERR_UNSUPPORTED;                 // the caller falls back to the dense 1-d copies
    }
    int rc = zero_gradients(a);
    if (rc != DPR_OK) return rc;
    if (a.P == 0 || a.B == 0) return launch_background_sum(a, grid.cells, dev);
    const float* pts = a.points;
    const float* pwt = a.point_weight;
    const int32_t* perm = nullptr;
    if (use_sort(a)) {
        const SortPlan sp = make_sort_plan(N_IN, a.P, 4, a.point_weight != nullptr, 256);
        rc = sort_points<float, N_IN>(a.points, a.point_weight, a.P, a.workspace, sp, dev, a.stream);
        if (rc != DPR_OK) return rc;
        char* ws = static_cast<char*>(a.workspace);
        pts = reinterpret_cast<const float*>(ws + sp.off_points);
        if (a.point_weight) pwt = reinterpret_cast<const float*>(ws + sp.off_pw);
        perm = reinterpret_cast<const int32_t*>(ws + sp.off_perm);
    }
    const int64_t ppc = (int64_t)kTmaConsumers * K;
    const int64_t point_chunks = (a.P + ppc - 1) / ppc;
    // pose chunks: fill whole waves of one CTA per SM, keep chunks long enough to amortise the per-point REDGs
    int64_t pose_chunks = 1;
    if (tuning().pose_chunk > 0) {
        pose_chunks = (a.B + tuning().pose_chunk - 1) / tuning().pose_chunk;
    } else {
        double best = 1e30;
        for (int64_t m = 1; m <= 16 && m <= a.B; ++m) {
            if (a.B / m < 32 && m > 1) break;
            const int64_t n = point_chunks * m;
            const int64_t waves = (n + dev.sm_count - 1) / dev.sm_count;
            const double cost = (double)waves * dev.sm_count / (double)n + 0.01 * m;   // wave inefficiency + REDG overhead
            if (cost < best) { best = cost; pose_chunks = m; }
        }
    }
    const int64_t pose_chunk = (a.B + pose_chunks - 1) / pose_chunks;
    pose_chunks = (a.B + pose_chunk - 1) / pose_chunk;
    if (point_chunks * pose_chunks > (int64_t)0x7fffffff) return DPR_ERR_BAD_DIMS;
    const size_t smem = tma_pullback_smem(tma2d_stage_words(a.grid, padded), stages, N_IN);
    const bool has_pw = a.point_weight != nullptr;
    auto launch = [&](auto kern) -> int {
        { const int rcs = opt_in_smem_once(kern, smem, dev); if (rcs != DPR_OK) return rcs; }
        LaunchScope scope("pullback_tma2d", a.stream);
        kern<<<(unsigned)(point_chunks * pose_chunks), kTmaConsumers + 32, smem, a.stream>>>(
            map, a.ds_dout, pts, a.rotation, a.translation, a.out_weight, pwt, a.d_points, a.d_rotation, a.d_translation,
            a.d_background, a.d_out_weight, a.d_point_weight, perm, grid, (int)a.P, a.B, (int)point_chunks, (int)pose_chunk);
        return DPR_OK;
    };
    if (padded) {
        if (stages == 3) rc = has_pw ? launch(pullback_tma2d_kernel<N_IN, K, true, 3, true>) : launch(pullback_tma2d_kernel<N_IN, K, false, 3, true>);
        else rc = has_pw ? launch(pullback_tma2d_kernel<N_IN, K, true, 2, true>) : launch(pullback_tma2d_kernel<N_IN, K, false, 2, true>);
    } else {
        if (stages == 3) rc = has_pw ? launch(pullback_tma2d_kernel<N_IN, K, true, 3, false>) : launch(pullback_tma2d_kernel<N_IN, K, false, 3, false>);
        else rc = has_pw ? launch(pullback_tma2d_kernel<N_IN, K, true, 2, false>) : launch(pullback_tma2d_kernel<N_IN, K, false, 2, false>);
    }
    if (rc != DPR_OK) return rc;
    DPR_CUDA_TRY(cudaGetLastError());
    set_last_path(DPR_OP_PULLBACK, padded ? (perm ? "tma2d_padded_sorted" : "tma2d_padded") : (perm ? "tma2d_sorted" : "tma2d"));
    return DPR_OK;
}

template <typename T>
int pullback_dispatch(const PullbackArgs<T>& a, const DeviceInfo& dev) {
    if (a.n_out == 3 && a.n_in == 3 && tuning().pullback_algo != 1 && a.P > 0 && a.B > 0 &&
        (tuning().pullback_algo == 7 || (tuning().pullback_algo == 0 && t3::worthwhile(a.grid, a.P, a.B)))) {
        const t3::Plan pl = t3::make_plan(a.n_in, a.grid, a.P, a.B, (int)sizeof(T));
        if (pl.ok && a.workspace && a.workspace_bytes >= pl.total) return pullback_tile3d<T, 3>(a, dev, pl);
    }
    // every other path may overwrite a workspace that holds cached bins of the 3-d tile path: un-mark them first
    if (tuning().binning_cache == 1 && a.workspace && a.workspace_bytes >= 256)
        DPR_CUDA_TRY(cudaMemsetAsync(static_cast<char*>(a.workspace) + t3::cache_valid_offset(), 0, sizeof(unsigned long long), a.stream));
    if constexpr (sizeof(T) == 4) {
        // TMA-staged kernel: whole pose image on chip, 16-byte aligned images, enough points to fill the CTAs
        const int64_t cells2 = a.n_out == 2 ? a.grid[0] * a.grid[1] : 0;
        const int64_t algo = tuning().pullback_algo;
        if (a.n_out == 2 && (a.n_in == 2 || a.n_in == 3) && (algo == 0 || algo == 4) && cells2 > 0 && (cells2 % 4) == 0 &&
            (reinterpret_cast<uintptr_t>(a.ds_dout) % 16) == 0 && a.P < (int64_t)0x3fffffff) {
            const bool worth = algo == 4 || (a.P >= 4 * (int64_t)kTmaConsumers * 8 && a.B >= 16);
            // padded rows first (tensor-map TMA, bank-skewed pitch), then the dense 1-d copies
            for (int padded = (tma2d_can_pad(a.grid) && tuning().tile3d_tma != 1) ? 1 : 0; worth && padded >= 0; --padded) {
                const int64_t words = tma2d_stage_words(a.grid, padded != 0);
                int stages = 0;
                if (tma_pullback_smem(words, 3, a.n_in) <= (size_t)dev.max_smem_optin - 1024) stages = 3;
                else if (tma_pullback_smem(words, 2, a.n_in) <= (size_t)dev.max_smem_optin - 1024) stages = 2;
                if (!stages) continue;
                const int rc = a.n_in == 2 ? pullback_tma2d<2>(a, dev, stages, padded != 0) : pullback_tma2d<3>(a, dev, stages, padded != 0);
                if (rc != DPR_ERR_UNSUPPORTED) return rc;
            }
        }
    }
    if (a.n_out == 2 && tuning().pullback_algo != 1 && a.P < (int64_t)0x3fffffff && a.grid[0] * a.grid[1] < (int64_t)0x3fffffff) {
        if (a.n_in == 2) return pullback_gather2d<T, 2>(a, dev);
        if (a.n_in == 3) return pullback_gather2d<T, 3>(a, dev);
    }
    // every (N_in, N_out) up to kMaxDim, like the reference's generated pullback (src/raster_pullback.jl:2-82)
#define DPR_PB_CASE(NI, NO) if (a.n_in == NI && a.n_out == NO) return pullback_global<T, NI, NO>(a, dev);
#define DPR_PB_ROW(NI) DPR_PB_CASE(NI, 1) DPR_PB_CASE(NI, 2) DPR_PB_CASE(NI, 3) DPR_PB_CASE(NI, 4)
    DPR_PB_ROW(1) DPR_PB_ROW(2) DPR_PB_ROW(3) DPR_PB_ROW(4)
#undef DPR_PB_ROW
#undef DPR_PB_CASE
    return DPR_ERR_UNSUPPORTED;
}

template int pullback_dispatch<float>(const PullbackArgs<float>&, const DeviceInfo&);
template int pullback_dispatch<double>(const PullbackArgs<double>&, const DeviceInfo&);

size_t pullback_workspace_bytes(int n_in, int n_out, const int64_t* grid, int64_t P, int64_t B, int sizeof_T) {
    const size_t sorted = make_sort_plan(n_in, P, sizeof_T, true, 256).total;
    size_t need = sorted;
    if (n_out == 3 && n_in == 3 && grid) {     // tile-binned 3-d path
        const t3::Plan pl = t3::make_plan(n_in, grid, P, B, sizeof_T);
        if (pl.ok && (tuning().pullback_algo == 7 || t3::worthwhile(grid, P, B)) && pl.total > need) need = pl.total;
    }
    return need;
}

}  // namespace dpr
