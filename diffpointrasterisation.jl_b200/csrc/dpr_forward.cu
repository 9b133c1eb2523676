// dpr_forward.cu - batched forward splat for sm_100a.
//
// Replaces the reference's canonical raster! (src/raster.jl:5-34) and raster_kernel! (src/raster.jl:36-66):
// rotate + translate + project each point, then scatter-add Prod(deltas) * out_weight * point_weight into the
// 2^N_out cells around it, on top of the per-pose background (src/raster.jl:27).
//
// Two kernel paths (measured ceilings in profiles/probe_atomics_r01.json):
//   tile   (2-d grids)  one CTA per (pose, slab of rows, point split) accumulates its slab in shared memory and
//                       flushes it once with coalesced 16-byte stores that also add the background - no init pass,
//                       no global atomics.  When the image is slightly larger than shared memory the CTA keeps the
//                       central band of rows on chip and sends the few border splats to L2 with REDG.
//   global (any grid)   one thread per (point, pose); native REDG.ADD.F32x2 on the two x-adjacent corners when
//                       8-byte aligned; the pose image stays L2-resident because CTAs are ordered pose-major.
#include "dpr_common.cuh"
#include "dpr_internal.h"

namespace dpr {

// ---------------------------------------------------------------------------------------------------------
// out[:, b] = background[b]        (src/raster.jl:27) - only used by the paths that accumulate with REDG
// ---------------------------------------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void __launch_bounds__(256) fill_background_kernel(T* __restrict__ out, const T* __restrict__ background,
                                                              int64_t cells, int64_t total_vec) {
    struct alignas(sizeof(T) * VEC) Pack { T v[VEC]; };
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += stride) {
        const int64_t e = i * VEC;
        Pack pk;
        if (VEC == 1 || (cells % VEC) == 0) {
            const T v = background ? __ldg(background + e / cells) : T(0);
#pragma unroll
            for (int k = 0; k < VEC; ++k) pk.v[k] = v;
        } else {
#pragma unroll
            for (int k = 0; k < VEC; ++k) pk.v[k] = background ? __ldg(background + (e + k) / cells) : T(0);
        }
        *reinterpret_cast<Pack*>(out + e) = pk;
    }
}

template <typename T>
static int launch_fill_background(T* out, const T* background, int64_t cells, int64_t B, const DeviceInfo& dev,
                                  cudaStream_t stream) {
    const int64_t total = cells * B;
    if (total == 0) return DPR_OK;
    if (!background) {
        DPR_CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(T) * (size_t)total, stream));
        return DPR_OK;
    }
    constexpr int VEC = 16 / sizeof(T);
    const bool vec_ok = (total % VEC) == 0 && (reinterpret_cast<uintptr_t>(out) % 16) == 0;
    const int64_t n = vec_ok ? total / VEC : total;
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)dev.sm_count * 32;
    if (blocks > cap) blocks = cap;
    {
        LaunchScope scope("fill_background", stream);
        if (vec_ok) fill_background_kernel<T, VEC><<<(unsigned)blocks, 256, 0, stream>>>(out, background, cells, n);
        else fill_background_kernel<T, 1><<<(unsigned)blocks, 256, 0, stream>>>(out, background, cells, n);
    }
    DPR_CUDA_TRY(cudaGetLastError());
    return DPR_OK;
}

// ---------------------------------------------------------------------------------------------------------
// global path: thread per (point, pose), REDG into the (pre-initialised) pose image
// ---------------------------------------------------------------------------------------------------------
template <typename T, int N_IN, int N_OUT>
__global__ void __launch_bounds__(256)
fwd_splat_global_kernel(const T* __restrict__ points, const T* __restrict__ rotation, const T* __restrict__ translation,
                        const T* __restrict__ out_weight, const T* __restrict__ point_weight, T* __restrict__ out,
                        Grid<T, N_OUT> grid, int64_t P, int chunks, int pts_per_cta) {
    const int64_t b = blockIdx.x / chunks;
    const int chunk = blockIdx.x % chunks;
    Pose<T, N_IN, N_OUT> pose;
    load_pose(pose, rotation, translation, out_weight, b);
    T* __restrict__ img = out + b * grid.cells;
    const int64_t p_begin = (int64_t)chunk * pts_per_cta;
    const int64_t p_end = (p_begin + pts_per_cta < P) ? p_begin + pts_per_cta : P;
    for (int64_t p = p_begin + threadIdx.x; p < p_end; p += blockDim.x) {
        T x[N_IN];
        load_point(x, points, p);
        int i0[N_OUT];
        T dl[N_OUT], du[N_OUT];
        if (!stencil(x, pose, grid, i0, dl)) continue;
        const T weight = pose.ow * (point_weight ? __ldg(point_weight + p) : T(1));  // src/raster.jl:51
#pragma unroll
        for (int k = 0; k < N_OUT; ++k) du[k] = T(1) - dl[k];
        const bool x_lo = i0[0] >= 0, x_hi = i0[0] + 1 < grid.g[0];
#pragma unroll
        for (int ch = 0; ch < (1 << (N_OUT - 1)); ++ch) {  // corners of the dimensions above the first
            bool inb = true;
            int64_t off = 0, stride = grid.g[0];
#pragma unroll
            for (int k = 1; k < N_OUT; ++k) {
                const int idx = i0[k] + ((ch >> (k - 1)) & 1);
                inb = inb && idx >= 0 && idx < grid.g[k];   // per-corner bounds rule, src/raster.jl:62
                off += idx * stride;
                stride *= grid.g[k];
            }
            if (!inb) continue;
            const T v0 = corner_weight<T, N_OUT>(ch << 1, dl, du) * weight;        // src/raster.jl:63
            const T v1 = corner_weight<T, N_OUT>((ch << 1) | 1, dl, du) * weight;
            T* addr = img + off + i0[0];
            if (x_lo && x_hi) red_add2(addr, v0, v1);
            else if (x_lo) red_add(addr, v0);
            else if (x_hi) red_add(addr + 1, v1);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// tile path (2-d): CTA = (pose b, slab s, point split q); slab rows [ys, ye) of the image live in shared memory.
// Rows outside the band [band_lo, band_hi) (hybrid mode, one slab) are accumulated in L2 with REDG.
// ---------------------------------------------------------------------------------------------------------
template <typename T>
struct TileParams {
    int slabs;        // S
    int splits;       // Q
    int rows;         // rows per slab (last slab may be shorter)
    int band_lo, band_hi;  // rows covered by shared-memory slabs
    int exclusive;    // 1: CTA owns its cells -> flush = store(tile + bg) and it initialises border rows itself
                      // 0: out was pre-filled with the background -> flush = REDG
};

template <typename T, int N_IN>
__global__ void __launch_bounds__(1024)
fwd_splat_tile2d_kernel(const T* __restrict__ points, const T* __restrict__ rotation, const T* __restrict__ translation,
                        const T* __restrict__ background, const T* __restrict__ out_weight,
                        const T* __restrict__ point_weight, T* __restrict__ out, Grid<T, 2> grid, int64_t P,
                        TileParams<T> tp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* tile = reinterpret_cast<T*>(smem_raw);
    const int per_pose = tp.slabs * tp.splits;
    const int64_t b = blockIdx.x / per_pose;
    const int rem = blockIdx.x % per_pose;
    const int s = rem / tp.splits, q = rem % tp.splits;
    const int g0 = grid.g[0], g1 = grid.g[1];
    const int ys = tp.band_lo + s * tp.rows;
    const int ye = (ys + tp.rows < tp.band_hi) ? ys + tp.rows : tp.band_hi;
    const int n_tile = (ye - ys) * g0;
    T* __restrict__ img = out + b * grid.cells;
    const T bg = background ? __ldg(background + b) : T(0);
    const bool border = (tp.band_lo > 0 || tp.band_hi < g1);

    for (int i = threadIdx.x; i < n_tile; i += blockDim.x) tile[i] = T(0);
    if (tp.exclusive && border) {
        // this CTA owns the whole pose image (hybrid => one slab, one split): background for the border rows
        const int lo_cells = tp.band_lo * g0;
        for (int i = threadIdx.x; i < lo_cells; i += blockDim.x) img[i] = bg;
        for (int i = tp.band_hi * g0 + threadIdx.x; i < g0 * g1; i += blockDim.x) img[i] = bg;
        __threadfence();  // the plain stores must reach L2 before this CTA's REDGs to the same cells
    }
    __syncthreads();

    Pose<T, N_IN, 2> pose;
    load_pose(pose, rotation, translation, out_weight, b);
    const int64_t per_split = (P + tp.splits - 1) / tp.splits;
    const int64_t p_begin = (int64_t)q * per_split;
    const int64_t p_end = (p_begin + per_split < P) ? p_begin + per_split : P;
    const bool do_border = border && s == 0;
    for (int64_t p = p_begin + threadIdx.x; p < p_end; p += blockDim.x) {
        T x[N_IN];
        load_point(x, points, p);
        int i0[2];
        T dl[2], du[2];
        if (!stencil(x, pose, grid, i0, dl)) continue;
        const T weight = pose.ow * (point_weight ? __ldg(point_weight + p) : T(1));
        du[0] = T(1) - dl[0];
        du[1] = T(1) - dl[1];
        const bool x_lo = i0[0] >= 0, x_hi = i0[0] + 1 < g0;
#pragma unroll
        for (int cy = 0; cy < 2; ++cy) {
            const int iy = i0[1] + cy;
            if (iy < 0 || iy >= g1) continue;
            const T v0 = corner_weight<T, 2>(cy << 1, dl, du) * weight;
            const T v1 = corner_weight<T, 2>((cy << 1) | 1, dl, du) * weight;
            if (iy >= ys && iy < ye) {
                T* addr = tile + (iy - ys) * g0 + i0[0];
                if (x_lo) atomicAdd(addr, v0);
                if (x_hi) atomicAdd(addr + 1, v1);
            } else if (do_border && (iy < tp.band_lo || iy >= tp.band_hi)) {
                T* addr = img + (int64_t)iy * g0 + i0[0];
                if (x_lo && x_hi) red_add2(addr, v0, v1);
                else if (x_lo) red_add(addr, v0);
                else if (x_hi) red_add(addr + 1, v1);
            }
        }
    }
    __syncthreads();

    T* __restrict__ dst = img + (int64_t)ys * g0;
    if (tp.exclusive) {
        constexpr int VEC = 16 / sizeof(T);
        struct alignas(16) Pack { T v[VEC]; };
        if ((n_tile % VEC) == 0 && (reinterpret_cast<uintptr_t>(dst) % 16) == 0) {
            for (int i = threadIdx.x; i < n_tile / VEC; i += blockDim.x) {
                Pack pk = reinterpret_cast<const Pack*>(tile)[i];
#pragma unroll
                for (int k = 0; k < VEC; ++k) pk.v[k] += bg;
                reinterpret_cast<Pack*>(dst)[i] = pk;
            }
        } else {
            for (int i = threadIdx.x; i < n_tile; i += blockDim.x) dst[i] = tile[i] + bg;
        }
    } else {
        for (int i = threadIdx.x; i < n_tile; i += blockDim.x) {
            const T v = tile[i];
            if (v != T(0)) red_add(dst + i, v);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// host-side planning + dispatch
// ---------------------------------------------------------------------------------------------------------
template <typename T, int N_OUT>
static Grid<T, N_OUT> make_grid(const int64_t* g) {
    Grid<T, N_OUT> grid;
    grid.cells = 1;
    for (int k = 0; k < N_OUT; ++k) {
        grid.g[k] = (int)g[k];
        grid.scale[k] = T(g[k]) / T(2);  // src/raster.jl:25
        grid.cells *= g[k];
    }
    return grid;
}

template <typename T, int N_IN, int N_OUT>
static int forward_global(const ForwardArgs<T>& a, const DeviceInfo& dev) {
    const Grid<T, N_OUT> grid = make_grid<T, N_OUT>(a.grid);
    int rc = launch_fill_background(a.out, a.background, grid.cells, a.B, dev, a.stream);
    if (rc != DPR_OK) return rc;
    if (a.P == 0 || a.B == 0) return DPR_OK;
    int pts_per_cta = 1024;
    int64_t chunks = (a.P + pts_per_cta - 1) / pts_per_cta;
    // keep the 1-d grid below 2^31 and give few-pose problems enough CTAs
    while (chunks * a.B > (int64_t)0x7fffffff) { pts_per_cta *= 2; chunks = (a.P + pts_per_cta - 1) / pts_per_cta; }
    while (pts_per_cta > 256 && chunks * a.B < (int64_t)dev.sm_count * 16) {
        pts_per_cta /= 2;
        chunks = (a.P + pts_per_cta - 1) / pts_per_cta;
    }
    {
        LaunchScope scope("fwd_splat_global", a.stream);
        fwd_splat_global_kernel<T, N_IN, N_OUT><<<(unsigned)(chunks * a.B), 256, 0, a.stream>>>(
            a.points, a.rotation, a.translation, a.out_weight, a.point_weight, a.out, grid, a.P, (int)chunks, pts_per_cta);
    }
    DPR_CUDA_TRY(cudaGetLastError());
    set_last_path(DPR_OP_FORWARD, "global_redg");
    return DPR_OK;
}

// Decide whether (and how) the 2-d tile kernel applies. Returns false if the global path should be used.
template <typename T>
static bool plan_tile2d(const ForwardArgs<T>& a, const DeviceInfo& dev, TileParams<T>& tp, size_t& smem_bytes) {
    const int64_t g0 = a.grid[0], g1 = a.grid[1];
    int64_t budget = tuning().tile_smem_bytes > 0 ? tuning().tile_smem_bytes : (int64_t)dev.max_smem_optin - 1024;
    if (budget > (int64_t)dev.max_smem_optin) budget = dev.max_smem_optin;
    const int64_t row_bytes = g0 * (int64_t)sizeof(T);
    const int64_t rows_fit = budget / row_bytes;
    if (rows_fit < 8 || g0 * g1 > (int64_t)0x3fffffff) return false;
    tp.band_lo = 0;
    tp.band_hi = (int)g1;
    if (rows_fit >= g1) {
        tp.slabs = 1;
        tp.rows = (int)g1;
    } else if (rows_fit * 10 >= g1 * 8) {
        // hybrid: keep the central band on chip, border rows go to L2 (few splats land there)
        tp.slabs = 1;
        tp.rows = (int)rows_fit;
        tp.band_lo = (int)((g1 - rows_fit) / 2);
        tp.band_hi = tp.band_lo + (int)rows_fit;
    } else {
        const int64_t S = (g1 + rows_fit - 1) / rows_fit;
        if (S > 8) return false;
        tp.slabs = (int)S;
        tp.rows = (int)((g1 + S - 1) / S);
    }
    // point splits: fill the machine when there are few (pose, slab) pairs
    int64_t Q = 1;
    const int64_t want = (int64_t)dev.sm_count * 2;
    if (tuning().point_split > 0) Q = tuning().point_split;
    else if (a.B * tp.slabs < want) {
        Q = (want + a.B * tp.slabs - 1) / (a.B * tp.slabs);
        const int64_t max_q = (a.P + 2047) / 2048;
        if (Q > max_q) Q = max_q;
        if (Q < 1) Q = 1;
    }
    if (a.B * tp.slabs * Q > (int64_t)0x7fffffff) return false;
    tp.splits = (int)Q;
    tp.exclusive = (Q == 1) ? 1 : 0;
    smem_bytes = (size_t)tp.rows * (size_t)row_bytes;
    return true;
}

template <typename T, int N_IN>
static int forward_tile2d(const ForwardArgs<T>& a, const DeviceInfo& dev, const TileParams<T>& tp, size_t smem_bytes) {
    const Grid<T, 2> grid = make_grid<T, 2>(a.grid);
    if (!tp.exclusive) {
        int rc = launch_fill_background(a.out, a.background, grid.cells, a.B, dev, a.stream);
        if (rc != DPR_OK) return rc;
    }
    auto kern = fwd_splat_tile2d_kernel<T, N_IN>;
    DPR_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    const int64_t ctas = a.B * tp.slabs * tp.splits;
    {
        LaunchScope scope("fwd_splat_tile2d", a.stream);
        kern<<<(unsigned)ctas, 1024, smem_bytes, a.stream>>>(a.points, a.rotation, a.translation, a.background,
                                                             a.out_weight, a.point_weight, a.out, grid, a.P, tp);
    }
    DPR_CUDA_TRY(cudaGetLastError());
    const bool border = tp.band_lo > 0 || tp.band_hi < (int)a.grid[1];
    set_last_path(DPR_OP_FORWARD, border ? "tile2d_hybrid" : (tp.slabs > 1 ? "tile2d_slabs" : (tp.exclusive ? "tile2d" : "tile2d_split")));
    return DPR_OK;
}

template <typename T>
int forward_dispatch(const ForwardArgs<T>& a, const DeviceInfo& dev) {
    const int64_t algo = tuning().forward_algo;
    if (a.n_out == 2 && algo != 1 && a.P > 0 && a.B > 0) {
        TileParams<T> tp;
        size_t smem = 0;
        if (plan_tile2d(a, dev, tp, smem)) {
            if (a.n_in == 2) return forward_tile2d<T, 2>(a, dev, tp, smem);
            if (a.n_in == 3) return forward_tile2d<T, 3>(a, dev, tp, smem);
        }
    }
    if (a.n_in == 2 && a.n_out == 2) return forward_global<T, 2, 2>(a, dev);
    if (a.n_in == 3 && a.n_out == 2) return forward_global<T, 3, 2>(a, dev);
    if (a.n_in == 3 && a.n_out == 3) return forward_global<T, 3, 3>(a, dev);
    return DPR_ERR_UNSUPPORTED;
}

template int forward_dispatch<float>(const ForwardArgs<float>&, const DeviceInfo&);
template int forward_dispatch<double>(const ForwardArgs<double>&, const DeviceInfo&);

size_t forward_workspace_bytes(int, int, const int64_t*, int64_t, int64_t, int) { return 0; }

}  // namespace dpr
