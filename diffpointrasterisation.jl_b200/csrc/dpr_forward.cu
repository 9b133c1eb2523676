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
#include "dpr_sort.cuh"
#include "dpr_tile3d.cuh"

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
//
// Two accumulation modes for the shared-memory tile:
//   float CAS    atomicAdd(float*) on shared memory = LDS + ATOMS.CAST.SPIN loop on sm_100a (no native f32 smem add)
//   fixed point  (Float32, non-negative weights) a contribution w * out_weight * point_weight is accumulated as the
//                integer rint(w * pw' * Q) (pw' = point weight scaled into [0, 1], 2^(F-1) <= Q <= 2^F, F <= 22) with the
//                NATIVE 32-bit ATOMS.ADD (3.4x the CAS rate, profiles/probe_atomics_r01.json).  The integer is the bit
//                pattern of a subnormal product - no conversion instruction (dpr_forward_fast.cuh).  Integer addition
//                is exact, so the only error is the quantisation (about 2^-F of the largest contribution per splat)
//                and the result is independent of the order of the atomics.  Wrap-around of a 32-bit cell is detected
//                exactly by a mass checksum (sum of all quantised contributions == sum of all cells); the CTA then
//                redoes its slab with the float CAS mode.
// ---------------------------------------------------------------------------------------------------------
template <typename T>
struct TileParams {
    int slabs;        // S
    int splits;       // Q
    int rows;         // rows per slab (last slab may be shorter)
    int band_lo, band_hi;  // rows covered by shared-memory slabs
    int exclusive;    // 1: CTA owns its cells -> flush = store(tile + bg) and it initialises border rows itself
                      // 0: out was pre-filled with the background -> flush = REDG
    int fixed_bits;   // planning only: fractional bits F for the Float32 fixed-point kernel (dpr_forward_fast.cuh)
};


// Generic slab accumulation with shared-memory atomicAdd in the element type (CAS loop on sm_100a): used for Float64
// and as the fallback of the Float32 fixed-point kernel.
// `load(p, x, pw)` fetches point p and its weight (AoS arrays here, the packed float4 copy in dpr_forward_radial.cuh).
template <typename T, int N_IN, typename Loader>
__device__ __forceinline__ void tile_accumulate_with(T* __restrict__ tile, T* __restrict__ img, Loader load,
                                                     const Pose<T, N_IN, 2>& pose, const Grid<T, 2>& grid, int p_begin,
                                                     int p_end, int ys, int ye, int band_lo, int band_hi, bool do_border,
                                                     int pitch = 0) {
    const int g0 = grid.g[0], g1 = grid.g[1];
    if (pitch == 0) pitch = g0;                  // words per tile row (a skewed pitch spreads compact blobs over banks)
    const int nrows = ye - ys;
    int p = p_begin + threadIdx.x;
    T xn[N_IN], pwn = T(1);
#pragma unroll
    for (int j = 0; j < N_IN; ++j) xn[j] = T(0);
    if (p < p_end) load(p, xn, pwn);
    while (p < p_end) {
        T x[N_IN];
#pragma unroll
        for (int j = 0; j < N_IN; ++j) x[j] = xn[j];
        const T pw = pwn;
        const int pn = p + blockDim.x;
        if (pn < p_end) load(pn, xn, pwn);   // prefetch the next point while this one is processed
        p = pn;
        int i0[2];
        T dl[2];
        if (!stencil(x, pose, grid, i0, dl)) continue;
        const T weight = pose.ow * pw;                                   // src/raster.jl:51
        const T du0 = T(1) - dl[0], du1 = T(1) - dl[1];
        const T v00 = (du0 * du1) * weight, v10 = (dl[0] * du1) * weight;   // src/raster.jl:63, 104-106
        const T v01 = (du0 * dl[1]) * weight, v11 = (dl[0] * dl[1]) * weight;
        const int ry = i0[1] - ys;
        auto tile_add = [&](int off, T v) { atomicAdd(tile + off, v); };
        if ((unsigned)i0[0] < (unsigned)(g0 - 1) && (unsigned)ry < (unsigned)(nrows - 1)) {
            // interior of the slab: all four corners are in bounds and on chip
            const int off = ry * pitch + i0[0];
            tile_add(off, v00);
            tile_add(off + 1, v10);
            tile_add(off + pitch, v01);
            tile_add(off + pitch + 1, v11);
        } else {
            const bool x_lo = i0[0] >= 0, x_hi = i0[0] + 1 < g0;
#pragma unroll
            for (int cy = 0; cy < 2; ++cy) {
                const int iy = i0[1] + cy;
                if (iy < 0 || iy >= g1) continue;                       // per-corner bounds rule, src/raster.jl:62
                const T va = cy ? v01 : v00, vb = cy ? v11 : v10;
                if (iy >= ys && iy < ye) {
                    const int off = (iy - ys) * pitch + i0[0];
                    if (x_lo) tile_add(off, va);
                    if (x_hi) tile_add(off + 1, vb);
                } else if (do_border && (iy < band_lo || iy >= band_hi)) {
                    T* addr = img + (int64_t)iy * g0 + i0[0];
                    if (x_lo && x_hi) red_add2(addr, va, vb);
                    else if (x_lo) red_add(addr, va);
                    else if (x_hi) red_add(addr + 1, vb);
                }
            }
        }
    }
}

template <typename T, int N_IN>
__device__ __forceinline__ void tile_accumulate(T* __restrict__ tile, T* __restrict__ img, const T* __restrict__ points,
                                                const T* __restrict__ point_weight, const Pose<T, N_IN, 2>& pose,
                                                const Grid<T, 2>& grid, int p_begin, int p_end, int ys, int ye,
                                                int band_lo, int band_hi, bool do_border, int pitch = 0) {
    auto load = [&](int p, T (&x)[N_IN], T& pw) {
        load_point(x, points, (int64_t)p);
        if (point_weight) pw = __ldg(point_weight + p);
    };
    tile_accumulate_with<T, N_IN>(tile, img, load, pose, grid, p_begin, p_end, ys, ye, band_lo, band_hi, do_border, pitch);
}

__device__ __forceinline__ long long block_sum_ll(long long v, long long* scratch) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    long long t = 0;
    const int nw = (blockDim.x + 31) >> 5;
    for (int i = 0; i < nw; ++i) t += scratch[i];
    return t;
}

template <typename T, int N_IN>
__global__ void __launch_bounds__(1024, 1)
fwd_splat_tile2d_kernel(const T* __restrict__ points, const T* __restrict__ rotation, const T* __restrict__ translation,
                        const T* __restrict__ background, const T* __restrict__ out_weight,
                        const T* __restrict__ point_weight, T* __restrict__ out, Grid<T, 2> grid, int P,
                        TileParams<T> tp) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
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

    Pose<T, N_IN, 2> pose;
    load_pose(pose, rotation, translation, out_weight, b);
    const int per_split = (P + tp.splits - 1) / tp.splits;
    const int p_begin = q * per_split;
    const int p_end = (p_begin + per_split < P) ? p_begin + per_split : P;
    const bool do_border = border && s == 0;

    for (int i = threadIdx.x; i < n_tile; i += blockDim.x) tile[i] = T(0);   // all-zero bits: 0.0f and integer 0
    if (tp.exclusive && border) {
        // this CTA owns the whole pose image (hybrid => one slab, one split): background for the border rows
        const int lo_cells = tp.band_lo * g0;
        for (int i = threadIdx.x; i < lo_cells; i += blockDim.x) img[i] = bg;
        for (int i = tp.band_hi * g0 + threadIdx.x; i < g0 * g1; i += blockDim.x) img[i] = bg;
        __threadfence();  // the plain stores must reach L2 before this CTA's REDGs to the same cells
    }
    __syncthreads();

    tile_accumulate<T, N_IN>(tile, img, points, point_weight, pose, grid, p_begin, p_end, ys, ye, tp.band_lo, tp.band_hi,
                             do_border);
    __syncthreads();

    auto cell_value = [&](int i) -> T { return tile[i]; };
    T* __restrict__ dst = img + (int64_t)ys * g0;
    if (tp.exclusive) {
        constexpr int VEC = 16 / sizeof(T);
        struct alignas(16) Pack { T v[VEC]; };
        if ((n_tile % VEC) == 0 && (reinterpret_cast<uintptr_t>(dst) % 16) == 0) {
            for (int i = threadIdx.x; i < n_tile / VEC; i += blockDim.x) {
                Pack pk;
#pragma unroll
                for (int k = 0; k < VEC; ++k) pk.v[k] = cell_value(i * VEC + k) + bg;
                reinterpret_cast<Pack*>(dst)[i] = pk;
            }
        } else {
            for (int i = threadIdx.x; i < n_tile; i += blockDim.x) dst[i] = cell_value(i) + bg;
        }
    } else {
        for (int i = threadIdx.x; i < n_tile; i += blockDim.x) {
            const T v = cell_value(i);
            if (v != T(0)) red_add(dst + i, v);
        }
    }
}

// {max, min, mean} of point_weight, for the fixed-point eligibility test: per-CTA partials, then a one-CTA finish
__global__ void __launch_bounds__(256) point_weight_stats_partial_kernel(const float* __restrict__ pw, int64_t P,
                                                                         float* __restrict__ partial /* [grid][4] */) {
    float mx = -3.4e38f, mn = 3.4e38f, sum = 0.f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    auto take = [&](float w) {
        mx = fmaxf(mx, w);
        mn = fminf(mn, w);
        sum += w;
        if (!(w == w)) mn = -1.f;    // NaN weights disable the fixed-point mode
    };
    // at most 15 CTAs (the partials share the 256-byte statistics block): 16-byte loads, four in flight per thread
    int64_t n_vec = 0;
    if ((reinterpret_cast<uintptr_t>(pw) & 15u) == 0) {
        n_vec = P / 4;
        const float4* __restrict__ pw4 = reinterpret_cast<const float4*>(pw);
        for (int64_t i = tid; i < n_vec; i += 4 * stride) {
            float4 q[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) q[u] = i + u * stride < n_vec ? __ldg(pw4 + i + u * stride) : make_float4(q[0].x, q[0].x, q[0].x, q[0].x);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (i + u * stride >= n_vec) break;
                take(q[u].x); take(q[u].y); take(q[u].z); take(q[u].w);
            }
        }
    }
    for (int64_t i = n_vec * 4 + tid; i < P; i += stride) take(__ldg(pw + i));
    __shared__ float smx[8], smn[8], ssum[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
    }
    if ((threadIdx.x & 31) == 0) { smx[threadIdx.x >> 5] = mx; smn[threadIdx.x >> 5] = mn; ssum[threadIdx.x >> 5] = sum; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) { mx = fmaxf(mx, smx[i]); mn = fminf(mn, smn[i]); sum += ssum[i]; }
        partial[blockIdx.x * 4] = mx; partial[blockIdx.x * 4 + 1] = mn; partial[blockIdx.x * 4 + 2] = sum;
    }
}
__global__ void __launch_bounds__(32) point_weight_stats_finish_kernel(const float* __restrict__ partial, int n, int64_t P,
                                                                       float* __restrict__ stats) {
    float mx = -3.4e38f, mn = 3.4e38f;
    double sum = 0.0;
    for (int i = threadIdx.x; i < n; i += 32) {
        mx = fmaxf(mx, partial[i * 4]);
        mn = fminf(mn, partial[i * 4 + 1]);
        sum += (double)partial[i * 4 + 2];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
    }
    if (threadIdx.x == 0) { stats[0] = mx; stats[1] = mn; stats[2] = (float)(sum / (double)(P > 0 ? P : 1)); }
}

}  // namespace dpr
#include "dpr_forward_fast.cuh"
#include "dpr_forward_radial.cuh"
namespace dpr {

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
    int rc = DPR_OK;
    if (a.P == 0 || a.B == 0) return launch_fill_background(a.out, a.background, grid.cells, a.B, dev, a.stream);
    // Spatially sorted points make the lanes of a warp hit neighbouring cells, so their REDGs share L2 sectors
    // (profiles/probe_atomics_r01.json: 0.34 T/s coherent vs 0.19 T/s random); the forward image does not depend
    // on the order of the points, so no permutation has to be undone.
    const T* pts = a.points;
    const T* pwt = a.point_weight;
    bool sorted = false;
    {
        const SortPlan sp = make_sort_plan(N_IN, a.P, (int)sizeof(T), a.point_weight != nullptr, 256);
        const bool want = tuning().point_sort == 1 || (tuning().point_sort == 0 && a.P >= 65536 && a.B >= 4);
        if (want && a.P < (int64_t)0x7fffffff && a.workspace && a.workspace_bytes >= sp.total) {
            rc = sort_points<T, N_IN>(a.points, a.point_weight, a.P, a.workspace, sp, dev, a.stream);
            if (rc != DPR_OK) return rc;
            char* ws = static_cast<char*>(a.workspace);
            pts = reinterpret_cast<const T*>(ws + sp.off_points);
            if (a.point_weight) pwt = reinterpret_cast<const T*>(ws + sp.off_pw);
            sorted = true;
        }
    }
    int pts_per_cta = 1024;
    int64_t chunks = (a.P + pts_per_cta - 1) / pts_per_cta;
    // Large images (3-d grids): fill and splat a group of poses at a time so the group (<= 48 MB) is still in the
    // 126 MB L2 when the REDGs arrive - otherwise every image goes to HBM after the fill and comes back for the adds.
    const int64_t img_bytes = grid.cells * (int64_t)sizeof(T);
    int64_t group = a.B;
    if (img_bytes * a.B > ((int64_t)96 << 20)) {
        group = ((int64_t)48 << 20) / img_bytes;
        if (group < 1) group = 1;
    }
    if ((a.B + group - 1) / group > 256) group = a.B;       // too many launches: one pass
    const bool grouped = group < a.B;
    if (!grouped) {
        rc = launch_fill_background(a.out, a.background, grid.cells, a.B, dev, a.stream);
        if (rc != DPR_OK) return rc;
    }
    for (int64_t b0 = 0; b0 < a.B; b0 += group) {
        const int64_t nb = (b0 + group < a.B) ? group : a.B - b0;
        if (grouped) {
            rc = launch_fill_background(a.out + b0 * grid.cells, a.background ? a.background + b0 : nullptr, grid.cells, nb, dev, a.stream);
            if (rc != DPR_OK) return rc;
        }
        int ppc = pts_per_cta;
        int64_t ch = chunks;
        // keep the 1-d grid below 2^31 and give few-pose problems enough CTAs
        while (ch * nb > (int64_t)0x7fffffff) { ppc *= 2; ch = (a.P + ppc - 1) / ppc; }
        while (ppc > 256 && ch * nb < (int64_t)dev.sm_count * 16) { ppc /= 2; ch = (a.P + ppc - 1) / ppc; }
        LaunchScope scope("fwd_splat_global", a.stream);
        fwd_splat_global_kernel<T, N_IN, N_OUT><<<(unsigned)(ch * nb), 256, 0, a.stream>>>(
            pts, a.rotation + b0 * (N_OUT * N_IN), a.translation + b0 * N_OUT, a.out_weight ? a.out_weight + b0 : nullptr, pwt,
            a.out + b0 * grid.cells, grid, a.P, (int)ch, ppc);
    }
    DPR_CUDA_TRY(cudaGetLastError());
    set_last_path(DPR_OP_FORWARD, sorted ? "global_redg_sorted" : "global_redg");
    return DPR_OK;
}

// Decide whether (and how) the 2-d tile kernel applies. Returns false if the global path should be used.
template <typename T>
static bool plan_tile2d(const ForwardArgs<T>& a, const DeviceInfo& dev, TileParams<T>& tp, size_t& smem_bytes, bool& use_fast) {
    const int64_t g0 = a.grid[0], g1 = a.grid[1];
    const int64_t opt_budget = tuning().tile_smem_bytes;
    int64_t budget = opt_budget > 0 ? opt_budget : (int64_t)dev.max_smem_optin - 1024;
    if (budget > (int64_t)dev.max_smem_optin - 1024) budget = (int64_t)dev.max_smem_optin - 1024;
    // Float32 takes the fixed-point kernel (dpr_forward_fast.cuh), which needs room for its per-warp queues and, with
    // point weights, 256 bytes of workspace for their statistics
    use_fast = sizeof(T) == 4 && tuning().forward_accum != 1 &&
               (!a.point_weight || (a.workspace && a.workspace_bytes >= 256));
    int64_t extra = use_fast ? (int64_t)fast_extra_smem(false) + 128 : 0;
    const int64_t extra_cull = use_fast ? (int64_t)fast_extra_smem(true) - (int64_t)fast_extra_smem(false) : 0;
    if (use_fast && tuning().tile_smem_bytes == 0) budget -= extra;
    const int64_t row_bytes = g0 * (int64_t)sizeof(T);
    const int64_t rows_fit = budget / row_bytes;
    if (rows_fit < 8 || g0 * g1 > (int64_t)0x3fffffff || a.P > (int64_t)0x3fffffff) return false;
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
        // several slabs: leave room for the surviving-chunk list of the culling kernel
        const int64_t rows_fit2 = (tuning().tile_smem_bytes == 0 ? budget - extra_cull : budget) / row_bytes;
        if (rows_fit2 < 8) return false;
        const int64_t S = (g1 + rows_fit2 - 1) / rows_fit2;
        if (S > 8) return false;
        tp.slabs = (int)S;
        tp.rows = (int)((g1 + S - 1) / S);
        extra += extra_cull;
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
    // fixed-point fractional bits: headroom for ~64x the mean number of points per cell before a 32-bit cell wraps
    tp.fixed_bits = 0;
    if (use_fast) {
        const double per_cell = (double)a.P / (double)(g0 * g1);
        int head = 6;
        while (head < 14 && (double)(1 << head) < 64.0 * per_cell + 64.0) ++head;
        tp.fixed_bits = 32 - head > 22 ? 22 : 32 - head;     // cells are unsigned 32-bit
    }
    smem_bytes = (size_t)tp.rows * (size_t)row_bytes;
    if (use_fast) smem_bytes = (smem_bytes + 127) / 128 * 128 + (size_t)extra;
    return true;
}

template <typename T, int N_IN>
static int forward_tile2d(const ForwardArgs<T>& a, const DeviceInfo& dev, const TileParams<T>& tp, size_t smem_bytes) {
    const Grid<T, 2> grid = make_grid<T, 2>(a.grid);
    if (!tp.exclusive) {
        int rc = launch_fill_background(a.out, a.background, grid.cells, a.B, dev, a.stream);
        if (rc != DPR_OK) return rc;
    }
    const TileParams<T>& tpl = tp;
    auto kern = fwd_splat_tile2d_kernel<T, N_IN>;
    { const int rcs = opt_in_smem_once(kern, smem_bytes, dev); if (rcs != DPR_OK) return rcs; }
    const int64_t ctas = a.B * tp.slabs * tp.splits;
    {
        LaunchScope scope("fwd_splat_tile2d", a.stream);
        kern<<<(unsigned)ctas, 1024, smem_bytes, a.stream>>>(a.points, a.rotation, a.translation, a.background,
                                                             a.out_weight, a.point_weight, a.out, grid, (int)a.P, tpl);
    }
    DPR_CUDA_TRY(cudaGetLastError());
    const bool border = tp.band_lo > 0 || tp.band_hi < (int)a.grid[1];
    set_last_path(DPR_OP_FORWARD, border ? "tile2d_hybrid" : (tp.slabs > 1 ? "tile2d_slabs" : (tp.exclusive ? "tile2d" : "tile2d_split")));
    return DPR_OK;
}

// Float32 fast path: packed FP32x2 stencil, fixed-point native shared-memory atomics, optional run culling
template <int N_IN>
static int forward_tile2d_fast(const ForwardArgs<float>& a, const DeviceInfo& dev, const TileParams<float>& tp, size_t smem_bytes) {
    const Grid<float, 2> grid = make_grid<float, 2>(a.grid);
    if (!tp.exclusive) {
        int rc = launch_fill_background(a.out, a.background, grid.cells, a.B, dev, a.stream);
        if (rc != DPR_OK) return rc;
    }
    FastTileParams fp;
    fp.slabs = tp.slabs; fp.splits = tp.splits; fp.rows = tp.rows; fp.band_lo = tp.band_lo; fp.band_hi = tp.band_hi;
    fp.exclusive = tp.exclusive; fp.fixed_bits = tp.fixed_bits; fp.pw_stats = nullptr; fp.aabb = nullptr;
    fp.pitch = (int)a.grid[0];
    const int64_t per_split = ((a.P + tp.splits - 1) / tp.splits + kChunk - 1) / kChunk * kChunk;
    fp.per_split = (int)per_split;
    const bool has_pw = a.point_weight != nullptr;
    if (has_pw) {
        // statistics live in the first 256 bytes of the workspace: [0..3] result, [4..] up to 15 CTA partials
        float* stats = static_cast<float*>(a.workspace);
        int n_part = (int)((a.P + 65535) / 65536);
        if (n_part > 15) n_part = 15;
        if (n_part < 1) n_part = 1;
        {
            LaunchScope scope("point_weight_stats", a.stream);
            point_weight_stats_partial_kernel<<<n_part, 256, 0, a.stream>>>(a.point_weight, a.P, stats + 4);
        }
        {
            LaunchScope scope("point_weight_stats_finish", a.stream);
            point_weight_stats_finish_kernel<<<1, 32, 0, a.stream>>>(stats + 4, n_part, a.P, stats);
        }
        fp.pw_stats = stats;
    }
    // One slab per pose and enough work to amortise a counting sort: radius-sorted points, straight-line loop for the
    // chunks that are interior for a pose by Cauchy-Schwarz (dpr_forward_radial.cuh)
    {
        const RadialPlan rp = make_radial_plan(a.P, 256);
        const int64_t ps = tuning().point_sort;
        const bool want = a.P >= 8192 && (ps == 1 || (ps == 0 && (double)a.P * (double)a.B >= 6.4e7));
        if (tp.slabs == 1 && want && a.workspace && a.workspace_bytes >= rp.total && a.P < (int64_t)0x3ffffc00) {
            int rc = radial_sort_points<N_IN>(a.points, a.point_weight, fp.pw_stats, a.P, a.workspace, rp, dev, a.stream);
            if (rc != DPR_OK) return rc;
            char* ws = static_cast<char*>(a.workspace);
            const float4* pts4 = reinterpret_cast<const float4*>(ws + rp.off_pts4);
            const float* rmax = reinterpret_cast<const float*>(ws + rp.off_rmax);
            const int64_t chunks_per_split = (rp.n_chunks + tp.splits - 1) / tp.splits;
            fp.per_split = (int)(chunks_per_split * kChunk);
            const int64_t ctas = a.B * tp.splits;
            auto launch = [&](auto kern) -> int {
                { const int rcs = opt_in_smem_once(kern, smem_bytes, dev); if (rcs != DPR_OK) return rcs; }
                LaunchScope scope("fwd_tile2d_radial", a.stream);
                kern<<<(unsigned)ctas, 1024, smem_bytes, a.stream>>>(pts4, rmax, a.rotation, a.translation, a.background,
                                                                     a.out_weight, a.out, grid, rp.n_chunks, fp);
                return DPR_OK;
            };
            rc = has_pw ? launch(fwd_tile2d_radial_kernel<N_IN, true>) : launch(fwd_tile2d_radial_kernel<N_IN, false>);
            if (rc != DPR_OK) return rc;
            DPR_CUDA_TRY(cudaGetLastError());
            const bool hybrid = tp.rows < (int)a.grid[1];
            set_last_path(DPR_OP_FORWARD, hybrid ? "tile2d_hybrid_radial_fixed" : (tp.exclusive ? "tile2d_radial_fixed" : "tile2d_split_radial_fixed"));
            return DPR_OK;
        }
    }
    // Several slabs per pose: sort the points spatially once and let every slab CTA skip the 1024-point runs whose
    // bounding box cannot reach its rows (otherwise each of the S slabs would transform all P points).
    const float* pts = a.points;
    const float* pwt = a.point_weight;
    const SortPlan sp = make_sort_plan(N_IN, a.P, 4, has_pw, 256);
    const int64_t n_runs = (a.P + kChunk - 1) / kChunk;
    const size_t aabb_bytes = sizeof(float) * 2 * N_IN * (size_t)n_runs;
    if (tp.slabs >= 2 && tuning().point_sort != 2 && a.workspace && a.workspace_bytes >= sp.total + aabb_bytes && a.B >= 2) {
        int rc = sort_points<float, N_IN>(a.points, a.point_weight, a.P, a.workspace, sp, dev, a.stream, /*spread=*/true);
        if (rc != DPR_OK) return rc;
        char* ws = static_cast<char*>(a.workspace);
        pts = reinterpret_cast<const float*>(ws + sp.off_points);
        if (has_pw) pwt = reinterpret_cast<const float*>(ws + sp.off_pw);
        float* aabb = reinterpret_cast<float*>(ws + sp.total);
        {
            LaunchScope scope("chunk_aabb", a.stream);
            chunk_aabb_kernel<float, N_IN><<<(unsigned)n_runs, 256, 0, a.stream>>>(pts, a.P, kChunk, aabb);
        }
        fp.aabb = aabb;
        // Sorted points put the 32 lanes of a warp into a blob a few pixels wide; with a row pitch that is a multiple of
        // the 32 banks the bank depends on the column only and such a blob serialises 4-8 ways (profiles/
        // probe_atomics_r01.json: coherent lanes 1.3 T/s at pitch 128 against 2.2 T/s at pitch 220).  Four padding words
        // per row make the bank (x + 4 y) mod 32.
        if ((a.grid[0] % 32) == 0) {
            const size_t padded = ((size_t)tp.rows * (size_t)(a.grid[0] + 4) * 4 + 127) / 128 * 128 + fast_extra_smem(true) + 128;
            if (padded + 1024 <= (size_t)dev.max_smem_optin) {      // same reserve as plan_tile2d: static smem of the kernel
                fp.pitch = (int)a.grid[0] + 4;
                smem_bytes = padded;
            }
        }
    }
    const int64_t ctas = a.B * tp.slabs * tp.splits;
    auto launch = [&](auto kern) -> int {
        { const int rcs = opt_in_smem_once(kern, smem_bytes, dev); if (rcs != DPR_OK) return rcs; }
        LaunchScope scope("fwd_tile2d_fast", a.stream);
        kern<<<(unsigned)ctas, 1024, smem_bytes, a.stream>>>(pts, a.rotation, a.translation, a.background,
                                                             a.out_weight, pwt, a.out, grid, (int)a.P, fp);
        return DPR_OK;
    };
    int rc;
    if (fp.aabb) rc = has_pw ? launch(fwd_tile2d_fast_kernel<N_IN, true, 2>) : launch(fwd_tile2d_fast_kernel<N_IN, false, 2>);
    else if (tp.slabs > 1) rc = has_pw ? launch(fwd_tile2d_fast_kernel<N_IN, true, 1>) : launch(fwd_tile2d_fast_kernel<N_IN, false, 1>);
    else rc = has_pw ? launch(fwd_tile2d_fast_kernel<N_IN, true, 0>) : launch(fwd_tile2d_fast_kernel<N_IN, false, 0>);
    if (rc != DPR_OK) return rc;
    DPR_CUDA_TRY(cudaGetLastError());
    const bool border = tp.band_lo > 0 || tp.band_hi < (int)a.grid[1];
    set_last_path(DPR_OP_FORWARD, border ? "tile2d_hybrid_fixed" : (tp.slabs > 1 ? (fp.aabb ? "tile2d_slabs_culled_fixed" : "tile2d_slabs_fixed") : (tp.exclusive ? "tile2d_fixed" : "tile2d_split_fixed")));
    return DPR_OK;
}

// 3-d grids: per-pose tile binning + one CTA per (pose, tile) accumulating in shared memory (dpr_tile3d.cuh)
template <typename T, int N_IN>
static int forward_tile3d(const ForwardArgs<T>& a, const DeviceInfo& dev, const t3::Plan& pl) {
    const Grid<T, 3> grid = t3::make_grid3<T>(a.grid);
    char* ws = static_cast<char*>(a.workspace);
    t3::CacheCtl cache;
    int rc = t3::cache_begin<T>(cache, ws, pl, N_IN, a.grid, a.points, a.point_weight, a.rotation, a.translation, a.P, a.B, dev, a.stream);
    if (rc != DPR_OK) return rc;
    rc = t3::presort<T, N_IN>(a.points, a.point_weight, a.P, ws, pl, cache, dev, a.stream);
    if (rc != DPR_OK) return rc;
    const size_t smem = sizeof(T) * (size_t)t3::FwdTile<T>::SIZE;
    auto kern = t3::fwd_tile3d_kernel<T, N_IN>;
    rc = opt_in_smem_once(kern, smem, dev);
    if (rc != DPR_OK) return rc;
    const t3::Pt4<T>* pts4 = reinterpret_cast<const t3::Pt4<T>*>(ws + pl.off_pts4);
    const uint32_t* cnt = reinterpret_cast<const uint32_t*>(ws + pl.tile_scan.off_data);
    const uint32_t* entries = reinterpret_cast<const uint32_t*>(ws + pl.off_entries);
    const uint32_t* pw_stats = a.point_weight ? reinterpret_cast<const uint32_t*>(ws + pl.sort_scan.off_ticket) : nullptr;
    // fixed-point fractional bits (Float32): headroom for ~64x the mean number of contributions per cell before a 32-bit
    // cell wraps (a wrap is detected exactly and the tile redone with float atomics); 0 = float atomics only
    int fixed_bits = 0;
    if (sizeof(T) == 4 && tuning().forward_accum != 1) {
        const double per_cell = 8.0 * (double)a.P / (double)grid.cells;
        int head = 6;
        while (head < 14 && (double)(1 << head) < 64.0 * per_cell + 64.0) ++head;
        fixed_bits = 32 - head > 22 ? 22 : 32 - head;
    }
    for (int64_t b0 = 0; b0 < a.B; b0 += pl.group) {
        const int64_t nb = (b0 + pl.group < a.B) ? pl.group : a.B - b0;
        rc = t3::bin_poses<T, N_IN>(a.rotation, a.translation, grid, a.P, b0, nb, ws, pl, cache, a.stream);
        if (rc != DPR_OK) return rc;
        LaunchScope scope("fwd_tile3d", a.stream);
        kern<<<dim3((unsigned)pl.tg.nt[0], (unsigned)pl.tg.nt[1], (unsigned)(pl.tg.nt[2] * nb)), t3::kThreads, smem, a.stream>>>(pts4, entries, cnt, a.rotation, a.translation, a.background,
                                                                              a.out_weight, a.out, grid, pl.tg, b0, pw_stats, a.P, fixed_bits, cache.valid);
    }
    DPR_CUDA_TRY(cudaGetLastError());
    set_last_path(DPR_OP_FORWARD, fixed_bits ? "tile3d_binned_fixed" : "tile3d_binned");
    return DPR_OK;
}

template <typename T>
int forward_dispatch(const ForwardArgs<T>& a, const DeviceInfo& dev) {
    const int64_t algo = tuning().forward_algo;
    if (a.n_out == 3 && a.n_in == 3 && algo != 1 && a.P > 0 && a.B > 0 && (algo == 3 || t3::worthwhile(a.grid, a.P, a.B))) {
        const t3::Plan pl = t3::make_plan(a.n_in, a.grid, a.P, a.B, (int)sizeof(T));
        if (pl.ok && a.workspace && a.workspace_bytes >= pl.total) return forward_tile3d<T, 3>(a, dev, pl);
    }
    // every other path may overwrite a workspace that holds cached bins of the 3-d tile path: un-mark them first
    if (tuning().binning_cache == 1 && a.workspace && a.workspace_bytes >= 256)
        DPR_CUDA_TRY(cudaMemsetAsync(static_cast<char*>(a.workspace) + t3::cache_valid_offset(), 0, sizeof(unsigned long long), a.stream));
    if (a.n_out == 2 && (a.n_in == 2 || a.n_in == 3) && algo != 1 && a.P > 0 && a.B > 0) {
        TileParams<T> tp;
        size_t smem = 0;
        bool use_fast = false;
        if (plan_tile2d(a, dev, tp, smem, use_fast)) {
            if constexpr (sizeof(T) == 4) {
                if (use_fast && a.n_in == 2) return forward_tile2d_fast<2>(a, dev, tp, smem);
                if (use_fast && a.n_in == 3) return forward_tile2d_fast<3>(a, dev, tp, smem);
            }
            if (a.n_in == 2) return forward_tile2d<T, 2>(a, dev, tp, smem);
            if (a.n_in == 3) return forward_tile2d<T, 3>(a, dev, tp, smem);
        }
    }
    // every (N_in, N_out) up to kMaxDim: the reference generates its kernel for any pair (src/raster.jl:36-66)
#define DPR_FWD_CASE(NI, NO) if (a.n_in == NI && a.n_out == NO) return forward_global<T, NI, NO>(a, dev);
#define DPR_FWD_ROW(NI) DPR_FWD_CASE(NI, 1) DPR_FWD_CASE(NI, 2) DPR_FWD_CASE(NI, 3) DPR_FWD_CASE(NI, 4)
    DPR_FWD_ROW(1) DPR_FWD_ROW(2) DPR_FWD_ROW(3) DPR_FWD_ROW(4)
#undef DPR_FWD_ROW
#undef DPR_FWD_CASE
    return DPR_ERR_UNSUPPORTED;
}

template int forward_dispatch<float>(const ForwardArgs<float>&, const DeviceInfo&);
template int forward_dispatch<double>(const ForwardArgs<double>&, const DeviceInfo&);

size_t forward_workspace_bytes(int n_in, int n_out, const int64_t* grid, int64_t P, int64_t B, int sizeof_T) {
    // point-weight statistics (256 B) + room for the spatially sorted copy of the points and their run boxes
    const SortPlan sp = make_sort_plan(n_in, P, sizeof_T, true, 256);
    const size_t morton = sp.total + sizeof(float) * 2 * (size_t)n_in * (size_t)((P + 1023) / 1024) + 256;
    const size_t radial = P >= 8192 ? make_radial_plan(P, 256).total + 256 : 0;   // radial kernel: large clouds only
    size_t need = morton > radial ? morton : radial;
    if (n_out == 3 && n_in == 3 && grid) {     // tile-binned 3-d path: sorted copy, keys, counters, P x poses entries
        const t3::Plan pl = t3::make_plan(n_in, grid, P, B, sizeof_T);
        if (pl.ok && (tuning().forward_algo == 3 || t3::worthwhile(grid, P, B)) && pl.total > need) need = pl.total;
    }
    return need;
}

}  // namespace dpr
