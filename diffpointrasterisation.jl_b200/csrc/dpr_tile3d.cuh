// dpr_tile3d.cuh - 3-d grids: per-pose tile binning + shared-memory tile kernels (forward and pullback).
//
// The north-star design for volumes (BASELINE.json): one CTA per (pose, output tile), the points
// binned into tiles per pose, splats accumulated in shared memory, every output cell written exactly ONCE with
// coalesced 16-byte stores that add the background - no init pass over `out` (src/raster.jl:27 is folded into the
// flush), no global atomics on `out`.  The pullback mirrors it: one CTA per (pose, tile) brings its ds_dout tile on chip
// with ONE read (d_background, src/raster_pullback.jl:78, is summed from the same staged tile), gathers from shared
// memory, and the pose-sum of d_points / d_point_weight is a 16-byte vector reduction into a packed L2-resident buffer.
//
// Binning without duplicates: a point's 2 x 2 x 2 stencil can straddle up to 8 tiles.  Instead of inserting it into 8
// lists (worst-case workspace 8 P B), every (point, pose) is stored ONCE, under
//     key = ((pose * n_tiles + home_tile) << 3) | pattern,
// where home_tile holds the lowest in-bounds corner and bit k of `pattern` says that the stencil also reaches the next
// tile along dimension k.  The CTA of tile t walks its own 8 sub-lists and, for each of its 7 lower neighbours t - d,
// the sub-lists whose pattern contains d (27 sub-lists in total), and touches only the corners inside its own tile;
// all contributions are linear in the corner values, so the pieces add up to the reference's result.  The workspace is
// exactly P * B entries.
//
// Steps per call (all on the caller's stream):
//   1. spatial pre-sort of the points (counting sort by Hilbert cell) into a packed {x, y, z, w} copy: a run of
//      consecutive points is then compact in space, hence compact in EVERY pose's volume, so the binning kernels below
//      aggregate their atomics per warp and write their entries in coalesced runs, and the tile kernels' point gathers
//      and vector reductions hit neighbouring addresses;
//   2. count (point, pose) pairs per key - warp-aggregated with match.any; 3. exclusive scan (single-pass, decoupled
//      look-back); 4. scatter the sorted point indices; 5. the tile kernel; 6. (pullback) un-permute the packed gradients.
#pragma once
#include <cuda.h>

#include <type_traits>

#include "dpr_common.cuh"
#include "dpr_internal.h"
#include "dpr_sort.cuh"

namespace dpr {
namespace t3 {

constexpr int TX = 32, TY = 16, TZ = 16;          // tile extent in cells (x contiguous: one 128-byte line per Float32 row)
constexpr int kTileCells = TX * TY * TZ;
constexpr int kThreads = 256;
constexpr uint32_t kNoKey = 0xffffffffu;

template <typename T>
struct alignas(4 * sizeof(T)) Pt4 { T x, y, z, w; };
constexpr int kStatMaxBits = 4, kStatBad = 5, kStatSum = 6;     // see "helpers shared by the tile kernels"

// ---------------------------------------------------------------------------------------------------------
// exclusive scan of n 32-bit counters in place, one pass (decoupled look-back, Merrill & Garland 2016): CTAs take their
// chunk from a ticket counter, publish {flag, value} in one 64-bit word and resolve their prefix with a warp-wide
// look-back.  `state` (one word per chunk) and `ticket` must be zero on entry.
// ---------------------------------------------------------------------------------------------------------
constexpr int kScanChunk = 4096;
#define DPR_SCAN_FLAG_AGGREGATE (1ull << 62)
#define DPR_SCAN_FLAG_PREFIX (2ull << 62)

static __global__ void __launch_bounds__(1024) scan_lookback_kernel(uint32_t* __restrict__ data, int64_t n,
                                                                    unsigned long long* __restrict__ state,
                                                                    uint32_t* __restrict__ ticket) {
    __shared__ uint32_t s_chunk, s_base, warp_tot[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_chunk = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t chunk = s_chunk;
    const int64_t i0 = (int64_t)chunk * kScanChunk + tid * 4;
    uint32_t v[4] = {0u, 0u, 0u, 0u};
    if (i0 + 3 < n) {
        const uint4 q = *reinterpret_cast<const uint4*>(data + i0);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) if (i0 + k < n) v[k] = data[i0 + k];
    }
    const uint32_t s = v[0] + v[1] + v[2] + v[3];
    uint32_t incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warp_tot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        warp_tot[lane] = w;
    }
    __syncthreads();
    const uint32_t block_excl = (warp ? warp_tot[warp - 1] : 0u) + incl - s;
    const uint32_t total = warp_tot[31];
    if (warp == 0) {
        volatile unsigned long long* st = state;
        uint32_t base = 0;
        if (chunk == 0) {
            if (lane == 0) st[0] = DPR_SCAN_FLAG_PREFIX | total;
        } else {
            if (lane == 0) st[chunk] = DPR_SCAN_FLAG_AGGREGATE | total;
            int64_t j = (int64_t)chunk - 1;
            while (true) {
                const int64_t idx = j - lane;
                unsigned long long sv = idx >= 0 ? st[idx] : DPR_SCAN_FLAG_PREFIX;
                while (__any_sync(0xffffffffu, (sv >> 62) == 0)) {
                    if ((sv >> 62) == 0) sv = st[idx];
                }
                const unsigned pmask = __ballot_sync(0xffffffffu, (sv >> 62) == 2);
                uint32_t val = (uint32_t)sv;
                if (pmask && lane > __ffs(pmask) - 1) val = 0;       // beyond the nearest full prefix
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
                base += val;
                if (pmask) break;
                j -= 32;
            }
            if (lane == 0) st[chunk] = DPR_SCAN_FLAG_PREFIX | (unsigned long long)(uint32_t)(base + total);
        }
        if (lane == 0) s_base = base;
    }
    __syncthreads();
    uint32_t run = s_base + block_excl;
    if (i0 + 3 < n) {
        uint4 o4;
        o4.x = run; o4.y = o4.x + v[0]; o4.z = o4.y + v[1]; o4.w = o4.z + v[2];
        *reinterpret_cast<uint4*>(data + i0) = o4;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (i0 + k < n) data[i0 + k] = run;
            run += v[k];
        }
    }
}

// region layout for one scan: [ticket (256 B)] [state: chunks x 8 B] [data: n x 4 B] - one memset clears all three
struct ScanRegion {
    size_t off_ticket = 0, off_state = 0, off_data = 0, bytes = 0;
    int64_t n = 0;
    int chunks = 0;
};
inline ScanRegion make_scan_region(size_t base, int64_t n) {
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    ScanRegion r;
    r.n = n;
    r.chunks = (int)((n + kScanChunk - 1) / kScanChunk);
    if (r.chunks < 1) r.chunks = 1;
    r.off_ticket = al(base);
    r.off_state = r.off_ticket + 256;
    r.off_data = al(r.off_state + sizeof(unsigned long long) * (size_t)r.chunks);
    r.bytes = al(r.off_data + sizeof(uint32_t) * (size_t)n) - r.off_ticket;
    return r;
}
static int clear_scan_region(char* ws, const ScanRegion& r, cudaStream_t stream) {
    DPR_CUDA_TRY(cudaMemsetAsync(ws + r.off_ticket, 0, r.bytes, stream));
    return DPR_OK;
}
// after the data was cleared by clear_scan_region and filled by a counting kernel
static int launch_scan(char* ws, const ScanRegion& r, cudaStream_t stream, const char* name) {
    LaunchScope scope(name, stream);
    scan_lookback_kernel<<<(unsigned)r.chunks, 1024, 0, stream>>>(reinterpret_cast<uint32_t*>(ws + r.off_data), r.n,
                                                                  reinterpret_cast<unsigned long long*>(ws + r.off_state),
                                                                  reinterpret_cast<uint32_t*>(ws + r.off_ticket));
    DPR_CUDA_TRY(cudaGetLastError());
    return DPR_OK;
}

// ---------------------------------------------------------------------------------------------------------
// 1. pre-sort: keys + histogram come from bin_count_kernel (dpr_sort.cuh); this scatter writes the packed copy
// ---------------------------------------------------------------------------------------------------------
template <typename T, int N_IN>
__global__ void __launch_bounds__(256) sort_scatter4_kernel(const T* __restrict__ points, const T* __restrict__ point_weight,
                                                            int64_t P, const uint32_t* __restrict__ keys,
                                                            uint32_t* __restrict__ offsets, int32_t* __restrict__ perm,
                                                            Pt4<T>* __restrict__ pts4, Pt4<T>* __restrict__ acc4,
                                                            uint32_t* __restrict__ stats) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    float wmax = 0.f, wsum = 0.f;
    bool bad = false;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += stride) {
        const uint32_t pos = atomicAdd(offsets + keys[p], 1u);
        Pt4<T> q;
        q.x = __ldg(points + p * N_IN);
        q.y = N_IN > 1 ? __ldg(points + p * N_IN + (N_IN > 1 ? 1 : 0)) : T(0);
        q.z = N_IN > 2 ? __ldg(points + p * N_IN + (N_IN > 2 ? 2 : 0)) : T(0);
        q.w = point_weight ? __ldg(point_weight + p) : T(1);
        pts4[pos] = q;
        perm[pos] = (int32_t)p;
        if (acc4) { Pt4<T> z; z.x = z.y = z.z = z.w = T(0); acc4[pos] = z; }
        const float w = (float)q.w;
        bad = bad || !(w >= 0.f) || !(w < 3e38f);     // negative, NaN or infinite weights: no fixed-point accumulation
        wmax = fmaxf(wmax, w);
        wsum += w;
    }
    if (stats && point_weight) {     // {max, any bad, sum} of the point weights (forward: fixed-point eligibility)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
            wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
        }
        const bool any_bad = __any_sync(0xffffffffu, bad);
        if ((threadIdx.x & 31) == 0) {
            if (any_bad) atomicOr(stats + kStatBad, 1u);
            else {
                atomicMax(stats + kStatMaxBits, __float_as_uint(wmax));     // non-negative floats order like their bit patterns
                atomicAdd(reinterpret_cast<float*>(stats + kStatSum), wsum);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// 2./4. per-pose tile keys: count, then scatter
// ---------------------------------------------------------------------------------------------------------
struct TileGeom {
    int nt[3];       // tiles per dimension
    int n_tiles;
};

template <typename T, int N_IN>
__device__ __forceinline__ void load_xyz(T (&x)[N_IN], const Pt4<T>& q) {
    x[0] = q.x;
    if constexpr (N_IN > 1) x[1] = q.y;
    if constexpr (N_IN > 2) x[2] = q.z;
}

// key of one (point, pose) pair from the lower-corner cell i0 (already validated by stencil(): -1 <= i0 <= g - 1)
__device__ __forceinline__ uint32_t tile_key(const int (&i0)[3], const int (&g)[3], const TileGeom& tg, int pose_local) {
    constexpr int TS[3] = {TX, TY, TZ};
    int home[3];
    unsigned pat = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const bool lo_ok = i0[k] >= 0, hi_ok = i0[k] + 1 < g[k];
        const int t_lo = (lo_ok ? i0[k] : 0) / TS[k], t_hi = (i0[k] + 1) / TS[k];
        home[k] = lo_ok ? t_lo : t_hi;
        if (lo_ok && hi_ok && t_hi != t_lo) pat |= 1u << k;
    }
    const uint32_t tile = (uint32_t)((home[2] * tg.nt[1] + home[1]) * tg.nt[0] + home[0]);
    return (((uint32_t)pose_local * (uint32_t)tg.n_tiles + tile) << 3) | pat;
}

template <typename T, int N_IN, bool SCATTER>
__global__ void __launch_bounds__(256) tile_bin_kernel(const Pt4<T>* __restrict__ pts4, int P, const T* __restrict__ rotation,
                                                       const T* __restrict__ translation, Grid<T, 3> grid, TileGeom tg,
                                                       uint32_t* __restrict__ cnt, uint32_t* __restrict__ entries, int64_t b0) {
    constexpr int K = 4;
    const int bl = blockIdx.y;
    Pose<T, N_IN, 3> pose;
    load_pose<T, N_IN, 3>(pose, rotation, translation, nullptr, b0 + bl);
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int p = (blockIdx.x * K + k) * 256 + (int)threadIdx.x;
        uint32_t key = kNoKey;
        if (p < P) {
            const Pt4<T> q = pts4[p];
            T x[N_IN];
            load_xyz<T, N_IN>(x, q);
            int i0[3];
            T dl[3];
            if (stencil<T, N_IN, 3>(x, pose, grid, i0, dl)) key = tile_key(i0, grid.g, tg, bl);
        }
        // spatially sorted points: the 32 lanes of a warp share a handful of keys -> one atomic per distinct key
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        const int leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if (key != kNoKey && lane == leader) {
            if (SCATTER) base = atomicAdd(cnt + key, (uint32_t)__popc(peers));
            else atomicAdd(cnt + key, (uint32_t)__popc(peers));
        }
        if (SCATTER) {
            base = __shfl_sync(0xffffffffu, base, leader);
            if (key != kNoKey) entries[base + __popc(peers & ((1u << lane) - 1u))] = (uint32_t)p;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// the 27 sub-lists a tile CTA has to walk: slot i in 0..63, d = i >> 3 (lower-neighbour offset, one bit per
// dimension), pat = i & 7; the slot is live when pat contains d and the neighbour exists
// ---------------------------------------------------------------------------------------------------------
struct Ranges {
    uint32_t start[64];
    uint32_t pref[65];
};
// two phases so that a persistent CTA can issue the loads for its NEXT tile before it processes the current one and
// consume them afterwards: warp 0, lane handles slots 2*lane and 2*lane+1
struct RangeLoad { uint32_t s[2], e[2]; };
__device__ __forceinline__ RangeLoad ranges_issue(const uint32_t* __restrict__ cnt, const TileGeom& tg, int bl, int tx, int ty, int tz) {
    RangeLoad rl;
    rl.s[0] = rl.s[1] = rl.e[0] = rl.e[1] = 0u;
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int slot = 2 * lane + h, d = slot >> 3, pat = slot & 7;
            const int nx = tx - (d & 1), ny = ty - ((d >> 1) & 1), nz = tz - ((d >> 2) & 1);
            if ((pat & d) == d && nx >= 0 && ny >= 0 && nz >= 0) {
                const uint32_t tile = (uint32_t)((nz * tg.nt[1] + ny) * tg.nt[0] + nx);
                const uint32_t key = (((uint32_t)bl * (uint32_t)tg.n_tiles + tile) << 3) | (uint32_t)pat;
                rl.e[h] = __ldg(cnt + key);                 // after the scatter pass cnt[key] is the END of list `key`
                rl.s[h] = key ? __ldg(cnt + key - 1) : 0u;
            }
        }
    }
    return rl;
}
__device__ __forceinline__ void ranges_finish(Ranges& r, const RangeLoad& rl) {
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        uint32_t len[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            r.start[2 * lane + h] = rl.s[h];
            len[h] = rl.e[h] - rl.s[h];
        }
        const uint32_t mine = len[0] + len[1];
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        r.pref[2 * lane] = incl - mine;
        r.pref[2 * lane + 1] = incl - mine + len[0];
        if (lane == 31) r.pref[64] = incl;
    }
}
__device__ __forceinline__ void build_ranges(Ranges& r, const uint32_t* __restrict__ cnt, const TileGeom& tg, int bl,
                                             int tx, int ty, int tz) {
    const RangeLoad rl = ranges_issue(cnt, tg, bl, tx, ty, tz);
    ranges_finish(r, rl);
}

// ---------------------------------------------------------------------------------------------------------
// helpers shared by the tile kernels
// ---------------------------------------------------------------------------------------------------------
// words of the 256-byte header of the pre-sort scan region (cleared with it): [0] scan ticket, then the point-weight
// statistics gathered by sort_scatter4_kernel for the fixed-point eligibility test of the forward kernel
// word of the tile scan region's header used as the persistent pullback kernel's work counter
constexpr int kWorkCounter = 8;

// Walks the (up to 27) sub-lists of a tile as one sequence, two entries ahead: index e of the concatenation -> sorted
// point index.  `slot` only moves forward because e grows.
struct EntryCursor {
    const Ranges* rg;
    const uint32_t* entries;
    int slot;
    __device__ __forceinline__ uint32_t fetch(uint32_t e) {
        while (e >= rg->pref[slot + 1]) ++slot;
        return __ldg(entries + rg->start[slot] + (e - rg->pref[slot]));
    }
};

__device__ __forceinline__ long long block_sum_i64(long long v, long long* scratch /* [kThreads / 32] */) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    long long t = 0;
#pragma unroll
    for (int i = 0; i < kThreads / 32; ++i) t += scratch[i];
    return t;
}

// ---------------------------------------------------------------------------------------------------------
// 5a. forward tile kernel: CTA = (pose, tile); accumulate in shared memory; one store per output cell
//
// Float32 accumulates in FIXED POINT on the native 32-bit shared-memory reduction (ATOMS.ADD without return: the float
// atomicAdd is an LDS + ATOMS.CAST.SPIN loop on sm_100a and made this kernel issue 470 instructions per warp-entry,
// profiles/ncu_r02_b_cfg3_summary.txt).  Same scheme as the 2-d kernels (dpr_forward_fast.cuh): a contribution
// w * out_weight * point_weight = w * pw' * A (pw' = point_weight * 2^-em in [0, 1), A = out_weight * 2^em) is added as the
// integer rint(w * pw' * Q), 2^(F-1) < Q <= 2^F <= 2^22, produced without a conversion (a product with the subnormal whose
// bit pattern is Q is rounded to exactly that integer); the flush multiplies by A / Q.  A wrapped cell is detected exactly
// by a mass checksum and the tile is then redone with float atomics, as it is for poses / weights that are not eligible
// (non-positive out_weight, negative or non-finite point weights, dynamic range above 64).
// ---------------------------------------------------------------------------------------------------------
template <typename T, int N_IN>
__global__ void __launch_bounds__(kThreads) fwd_tile3d_kernel(const Pt4<T>* __restrict__ pts4, const uint32_t* __restrict__ entries,
                                                              const uint32_t* __restrict__ cnt, const T* __restrict__ rotation,
                                                              const T* __restrict__ translation, const T* __restrict__ background,
                                                              const T* __restrict__ out_weight, T* __restrict__ out,
                                                              Grid<T, 3> grid, TileGeom tg, int64_t b0,
                                                              const uint32_t* __restrict__ pw_stats, int64_t P, int fixed_bits) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* tile = reinterpret_cast<T*>(smem_raw);
    __shared__ Ranges rg;
    __shared__ long long scratch[kThreads / 32];
    const int bl = blockIdx.x / tg.n_tiles;
    const int t = blockIdx.x % tg.n_tiles;
    const int tx = t % tg.nt[0], ty = (t / tg.nt[0]) % tg.nt[1], tz = t / (tg.nt[0] * tg.nt[1]);
    const int64_t b = b0 + bl;
    build_ranges(rg, cnt, tg, bl, tx, ty, tz);
    constexpr int VEC = 16 / sizeof(T);
    struct alignas(16) Pack { T v[VEC]; };
    const int ox = tx * TX, oy = ty * TY, oz = tz * TZ;
    const T bg = background ? __ldg(background + b) : T(0);
    T* __restrict__ img = out + b * grid.cells;
    const bool vec_ok = (grid.g[0] % VEC) == 0 && (reinterpret_cast<uintptr_t>(out) % 16) == 0;
    __syncthreads();
    const uint32_t total = rg.pref[64];

    // stores `value(cell) + bg` for every cell of the tile that lies inside the volume; returns this thread's part of
    // the integer cell sum (fixed-point mode)
    auto flush = [&](bool fixed, float inv_q, bool empty) -> long long {
        long long cells = 0;
        for (int i = threadIdx.x; i < kTileCells / VEC; i += kThreads) {
            const int x4 = (i % (TX / VEC)) * VEC, y = (i / (TX / VEC)) % TY, z = i / ((TX / VEC) * TY);
            const int gx = ox + x4, gy = oy + y, gz = oz + z;
            Pack pk;
            if (empty) {
#pragma unroll
                for (int k = 0; k < VEC; ++k) pk.v[k] = bg;
            } else {
                pk = reinterpret_cast<const Pack*>(tile)[i];
                if constexpr (sizeof(T) == 4) {
                    if (fixed) {
#pragma unroll
                        for (int k = 0; k < VEC; ++k) {
                            const uint32_t u = __float_as_uint(pk.v[k]);
                            cells += u;
                            pk.v[k] = (float)u * inv_q;
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < VEC; ++k) pk.v[k] += bg;
            }
            if (gx >= grid.g[0] || gy >= grid.g[1] || gz >= grid.g[2]) continue;
            T* dst = img + ((int64_t)gz * grid.g[1] + gy) * grid.g[0] + gx;
            if (vec_ok && gx + VEC <= grid.g[0]) {
                __stcs(reinterpret_cast<float4*>(dst), *reinterpret_cast<const float4*>(&pk));
            } else {
#pragma unroll
                for (int k = 0; k < VEC; ++k) if (gx + k < grid.g[0]) dst[k] = pk.v[k];
            }
        }
        return cells;
    };
    if (total == 0) {            // nothing lands here: the tile is the background (src/raster.jl:27)
        flush(false, 0.f, true);
        return;
    }
    auto zero_tile = [&]() {
        for (int i = threadIdx.x; i < kTileCells / VEC; i += kThreads) {
            Pack z;
#pragma unroll
            for (int k = 0; k < VEC; ++k) z.v[k] = T(0);
            reinterpret_cast<Pack*>(tile)[i] = z;
        }
    };
    zero_tile();
    Pose<T, N_IN, 3> pose;
    load_pose<T, N_IN, 3>(pose, rotation, translation, out_weight, b);

    // ---- fixed-point eligibility and scale (uniform over the CTA) ---------------------------------------------
    bool fixed = false;
    float wq_den = 0.f, inv_q = 0.f, pw_scale = 1.f;
    if constexpr (sizeof(T) == 4) {
        bool ok = pose.ow > 0.f && fixed_bits > 0;
        int em = 0;
        if (pw_stats) {
            const float wmax = __uint_as_float(__ldg(pw_stats + kStatMaxBits));
            const float wmean = __uint_as_float(__ldg(pw_stats + kStatSum)) / (float)P;
            ok = ok && __ldg(pw_stats + kStatBad) == 0u && wmax > 0.f && wmax < 3e38f && wmax <= 64.f * wmean;
            if (ok) frexpf(wmax, &em);
        }
        const float A = ldexpf((float)pose.ow, em);
        ok = ok && A > 1e-30f && A < 1e30f;
        if (ok) {
            int e;
            frexpf(A, &e);                                        // A < 2^e
            const float Q = rintf(ldexpf(A, fixed_bits - e));     // 2^(F-1) <= Q <= 2^F <= 2^22
            wq_den = __int_as_float((int)Q);
            inv_q = A / Q;
            pw_scale = ldexpf(1.f, -em);
            fixed = true;
        }
    }
    __syncthreads();

    long long mass = 0;
    // one pass over the tile's entries; FIXED selects the accumulation mode at compile time
    auto accumulate = [&](auto fixed_tag) {
        constexpr bool FIXED = decltype(fixed_tag)::value;
        EntryCursor cur{&rg, entries, 0};
        uint32_t e = threadIdx.x;
        uint32_t idx_a = e < total ? cur.fetch(e) : 0u;
        uint32_t idx_b = e + kThreads < total ? cur.fetch(e + kThreads) : 0u;
        Pt4<T> qn = pts4[idx_a];
        const uint32_t tile_s = smem_u32(tile);
        for (; e < total; e += kThreads) {
            const Pt4<T> q = qn;
            idx_a = idx_b;
            if (e + kThreads < total) qn = pts4[idx_a];
            if (e + 2 * kThreads < total) idx_b = cur.fetch(e + 2 * kThreads);
            T x[N_IN];
            load_xyz<T, N_IN>(x, q);
            int i0[3];
            T dl[3], du[3];
            if (!stencil<T, N_IN, 3>(x, pose, grid, i0, dl)) continue;      // cannot happen: binned with the same arithmetic
#pragma unroll
            for (int k = 0; k < 3; ++k) du[k] = T(1) - dl[k];
            const int lx = i0[0] - ox, ly = i0[1] - oy, lz = i0[2] - oz;
            // corner values in bit order (x = bit 0): w_c * out_weight * point_weight (src/raster.jl:51,63,104-106), or its
            // fixed-point image
            T v[8];
            if constexpr (FIXED) {
                const float wq = wq_den * ((float)q.w * pw_scale);
                const float az[2] = {(float)du[2] * wq, (float)dl[2] * wq};
#pragma unroll
                for (int c = 0; c < 8; ++c) v[c] = T(((c & 1) ? (float)dl[0] : (float)du[0]) * (((c & 2) ? (float)dl[1] : (float)du[1]) * az[c >> 2]));
            } else {
                const T weight = pose.ow * q.w;
#pragma unroll
                for (int c = 0; c < 8; ++c) v[c] = corner_weight<T, 3>(c, dl, du) * weight;
            }
            const bool interior = (unsigned)lx < (unsigned)(TX - 1) && (unsigned)ly < (unsigned)(TY - 1) && (unsigned)lz < (unsigned)(TZ - 1) &&
                                  i0[0] + 1 < grid.g[0] && i0[1] + 1 < grid.g[1] && i0[2] + 1 < grid.g[2];
            const int off = (lz * TY + ly) * TX + lx;
            if (interior) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int o = off + (c & 1) + ((c >> 1) & 1) * TX + ((c >> 2) & 1) * TX * TY;
                    if constexpr (FIXED) {
                        const uint32_t qv = __float_as_uint((float)v[c]);
                        asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(tile_s + (uint32_t)o * 4u), "r"(qv) : "memory");
                        mass += qv;
                    } else {
                        atomicAdd(tile + o, v[c]);
                    }
                }
            } else {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int cx = lx + (c & 1), cy = ly + ((c >> 1) & 1), cz = lz + ((c >> 2) & 1);
                    // inside this CTA's tile and inside the grid (per-corner bounds rule, src/raster.jl:62)
                    const bool in = (unsigned)cx < (unsigned)TX && (unsigned)cy < (unsigned)TY && (unsigned)cz < (unsigned)TZ &&
                                    ox + cx < grid.g[0] && oy + cy < grid.g[1] && oz + cz < grid.g[2];
                    if (!in) continue;
                    const int o = (cz * TY + cy) * TX + cx;
                    if constexpr (FIXED) {
                        const uint32_t qv = __float_as_uint((float)v[c]);
                        asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(tile_s + (uint32_t)o * 4u), "r"(qv) : "memory");
                        mass += qv;
                    } else {
                        atomicAdd(tile + o, v[c]);
                    }
                }
            }
        }
    };
    if constexpr (sizeof(T) == 4) {
        if (fixed) {
            accumulate(std::true_type{});
            __syncthreads();
            // optimistic flush; a wrapped 32-bit cell shows as (sum of cells) != (sum of contributions), exactly
            const long long cells = flush(true, inv_q, false);
            const long long diff = block_sum_i64(cells - mass, scratch);
            if (diff == 0) return;
            zero_tile();
            __syncthreads();
        }
    }
    accumulate(std::false_type{});
    __syncthreads();
    flush(false, 0.f, false);
}

// ---------------------------------------------------------------------------------------------------------
// 5b. pullback tile kernel
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_tile4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void red_add4(Pt4<float>* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add4(Pt4<double>* addr, double a, double b, double c, double d) {
    atomicAdd(&addr->x, a);
    atomicAdd(&addr->y, b);
    atomicAdd(&addr->z, c);
    atomicAdd(&addr->w, d);
}

template <typename T, int N_IN, bool USE_TMA>
__global__ void __launch_bounds__(kThreads)
pullback_tile3d_simple_kernel(const __grid_constant__ CUtensorMap map, const T* __restrict__ ds_dout, const Pt4<T>* __restrict__ pts4,
                       const uint32_t* __restrict__ entries, const uint32_t* __restrict__ cnt, const T* __restrict__ rotation,
                       const T* __restrict__ translation, const T* __restrict__ out_weight, Pt4<T>* __restrict__ acc4,
                       T* __restrict__ d_rotation, T* __restrict__ d_translation, T* __restrict__ d_background,
                       T* __restrict__ d_out_weight, Grid<T, 3> grid, TileGeom tg, int64_t b0) {
    constexpr int NR = 3 * N_IN, NV = NR + 3 + 1;     // d_rotation (col-major 3 x N_IN), d_translation, d_out_weight
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* tile = reinterpret_cast<T*>(smem_raw);
    __shared__ Ranges rg;
    __shared__ __align__(8) uint64_t bar;
    __shared__ T red[kThreads / 32][NV + 1];
    const int bl = blockIdx.x / tg.n_tiles;
    const int t = blockIdx.x % tg.n_tiles;
    const int tx = t % tg.nt[0], ty = (t / tg.nt[0]) % tg.nt[1], tz = t / (tg.nt[0] * tg.nt[1]);
    const int64_t b = b0 + bl;
    const int ox = tx * TX, oy = ty * TY, oz = tz * TZ;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int VEC = 16 / sizeof(T);
    struct alignas(16) Pack { T v[VEC]; };

    if (USE_TMA) {
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            mbar_fence_init();
            mbar_arrive_expect_tx(&bar, (uint32_t)(kTileCells * sizeof(T)));
            tma_load_tile4d(tile, &map, ox, oy, oz, (int)b, &bar);          // cells outside the volume arrive as zeros
        }
    } else {
        const T* __restrict__ img = ds_dout + b * grid.cells;
        const bool vec_ok = (grid.g[0] % VEC) == 0 && (reinterpret_cast<uintptr_t>(ds_dout) % 16) == 0;
#pragma unroll 4
        for (int i = threadIdx.x; i < kTileCells / VEC; i += kThreads) {
            const int x4 = (i % (TX / VEC)) * VEC, y = (i / (TX / VEC)) % TY, z = i / ((TX / VEC) * TY);
            const int gx = ox + x4, gy = oy + y, gz = oz + z;
            Pack pk;
#pragma unroll
            for (int k = 0; k < VEC; ++k) pk.v[k] = T(0);
            if (gx < grid.g[0] && gy < grid.g[1] && gz < grid.g[2]) {
                const T* src = img + ((int64_t)gz * grid.g[1] + gy) * grid.g[0] + gx;
                if (vec_ok && gx + VEC <= grid.g[0]) {
                    const float4 q = __ldcs(reinterpret_cast<const float4*>(src));
                    pk = *reinterpret_cast<const Pack*>(&q);
                } else {
#pragma unroll
                    for (int k = 0; k < VEC; ++k) if (gx + k < grid.g[0]) pk.v[k] = __ldg(src + k);
                }
            }
            reinterpret_cast<Pack*>(tile)[i] = pk;
        }
    }
    build_ranges(rg, cnt, tg, bl, tx, ty, tz);
    __syncthreads();                       // ranges (and the mbarrier init / the cooperative tile load) visible
    if (USE_TMA) mbar_wait(&bar, 0);

    T acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = T(0);
    T bg_part = T(0);
    if (d_background) {                    // src/raster_pullback.jl:78 - from the staged tile (zero outside the volume)
        for (int i = threadIdx.x; i < kTileCells / VEC; i += kThreads) {
            const Pack pk = reinterpret_cast<const Pack*>(tile)[i];
#pragma unroll
            for (int k = 0; k < VEC; ++k) bg_part += pk.v[k];
        }
    }
    const uint32_t total = rg.pref[64];
    if (total) {
        Pose<T, N_IN, 3> pose;
        load_pose<T, N_IN, 3>(pose, rotation, translation, out_weight, b);
        int slot = 0;
        for (uint32_t e = threadIdx.x; e < total; e += kThreads) {
            while (e >= rg.pref[slot + 1]) ++slot;
            const uint32_t idx = __ldg(entries + rg.start[slot] + (e - rg.pref[slot]));
            const Pt4<T> q = pts4[idx];
            T x[N_IN];
            load_xyz<T, N_IN>(x, q);
            int i0[3];
            T dl[3], du[3];
            if (!stencil<T, N_IN, 3>(x, pose, grid, i0, dl)) continue;
#pragma unroll
            for (int k = 0; k < 3; ++k) du[k] = T(1) - dl[k];
            const int lx = i0[0] - ox, ly = i0[1] - oy, lz = i0[2] - oz;
            T s = T(0), gk[3] = {T(0), T(0), T(0)};
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int cx = lx + (c & 1), cy = ly + ((c >> 1) & 1), cz = lz + ((c >> 2) & 1);
                // corners of other tiles are that tile's CTA's job; corners outside the volume read the zero fill, which
                // equals skipping them (src/raster_pullback.jl:51)
                const bool in = (unsigned)cx < (unsigned)TX && (unsigned)cy < (unsigned)TY && (unsigned)cz < (unsigned)TZ;
                const T G = in ? tile[(cz * TY + cy) * TX + cx] : T(0);
                s += corner_weight<T, 3>(c, dl, du) * G;                                        // :55-58
#pragma unroll
                for (int n = 0; n < 3; ++n) {                                                    // :60-65, :150-160
                    T iw = ((c >> n) & 1) ? T(1) : T(-1);
#pragma unroll
                    for (int m = 0; m < 3; ++m)
                        if (m != n) iw *= ((c >> m) & 1) ? dl[m] : du[m];
                    gk[n] += G * iw;
                }
            }
            acc[NV - 1] += s * q.w;                           // d_out_weight,   src/raster_pullback.jl:57
            const T f = pose.ow * q.w;                        // :60
            T scaled[3];
#pragma unroll
            for (int n = 0; n < 3; ++n) {
                scaled[n] = (f * gk[n]) * grid.scale[n];      // :67
                acc[NR + n] += scaled[n];                     // d_translation, :68
            }
            T dpt[3] = {T(0), T(0), T(0)};
#pragma unroll
            for (int j = 0; j < N_IN; ++j) {
                T d = T(0);
#pragma unroll
                for (int n = 0; n < 3; ++n) {
                    acc[n + j * 3] += scaled[n] * x[j];       // d_rotation, :69
                    d += pose.R[n][j] * scaled[n];            // R' * scaled, :70
                }
                dpt[j] = d;
            }
            // pose-sum of d_points (:71, :141) and d_point_weight (:58, :146): one 16-byte reduction into the packed,
            // L2-resident buffer (sorted point order)
            red_add4(acc4 + idx, dpt[0], dpt[1], dpt[2], s * pose.ow);
        }
    }
    // per-pose sums of this CTA: warp shuffles, then one line of shared memory per warp, then one REDG per value
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = warp_sum(acc[v]);
    bg_part = warp_sum(bg_part);
    if (lane == 0) {
#pragma unroll
        for (int v = 0; v < NV; ++v) red[warp][v] = acc[v];
        red[warp][NV] = bg_part;
    }
    __syncthreads();
    if (threadIdx.x <= NV) {
        const int v = threadIdx.x;
        T r = T(0);
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) r += red[w][v];
        if (v == NV) { if (d_background) red_add(d_background + b, r); }
        else if (r != T(0)) {
            if (v < NR) red_add(d_rotation + b * NR + v, r);
            else if (v < NR + 3) red_add(d_translation + b * 3 + (v - NR), r);
            else if (d_out_weight) red_add(d_out_weight + b, r);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// 5c. pullback tile kernel, persistent: the production path when rows are 16-byte multiples.
//
// The one-CTA-per-tile kernel above spends 37 % of its stall samples waiting for its own tile load (profiles/
// ncu_r02_b_cfg3_summary.txt: the loads are synchronous and only three CTAs fit an SM) and pays a block reduction per
// tile.  Here a CTA keeps pulling (pose, tile) tickets from a global counter; the ds_dout tile of the NEXT ticket is
// streamed into the other half of a two-stage shared-memory ring with 16-byte cp.async (LDGSTS, zero-filled outside the
// volume) and its sub-list ranges are fetched by warp 0 while the current tile is processed; tickets are taken two ahead
// so the atomic's latency is hidden too.  Tickets of one CTA increase, so the pose only moves forward: the 3 N_in + 5
// per-pose sums (and the d_background partial) stay in registers across tiles and are block-reduced only when the pose
// changes.  The sample and its three derivatives come from the factorised trilinear form (25 flops instead of ~150).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool valid) {
    const int bytes = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T, int N_IN>
__global__ void __launch_bounds__(kThreads, sizeof(T) == 4 ? 3 : 1)
pullback_tile3d_kernel(const T* __restrict__ ds_dout, const Pt4<T>* __restrict__ pts4, const uint32_t* __restrict__ entries,
                       const uint32_t* __restrict__ cnt, const T* __restrict__ rotation, const T* __restrict__ translation,
                       const T* __restrict__ out_weight, Pt4<T>* __restrict__ acc4, T* __restrict__ d_rotation,
                       T* __restrict__ d_translation, T* __restrict__ d_background, T* __restrict__ d_out_weight,
                       Grid<T, 3> grid, TileGeom tg, int64_t b0, int n_work, uint32_t* __restrict__ work_counter) {
    constexpr int NR = 3 * N_IN, NV = NR + 3 + 1;     // d_rotation (col-major 3 x N_IN), d_translation, d_out_weight
    constexpr int VEC = 16 / sizeof(T);
    struct alignas(16) Pack { T v[VEC]; };
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* const ring = reinterpret_cast<T*>(smem_raw);       // two stages of kTileCells
    __shared__ Ranges rg[2];
    __shared__ T red[kThreads / 32][NV + 1];
    __shared__ int s_ticket[4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    auto tile_coords = [&](int w, int& bl, int& tx, int& ty, int& tz) {
        bl = w / tg.n_tiles;
        const int t = w % tg.n_tiles;
        tx = t % tg.nt[0]; ty = (t / tg.nt[0]) % tg.nt[1]; tz = t / (tg.nt[0] * tg.nt[1]);
    };
    // asynchronous copy of work item w's ds_dout tile into ring stage st (16-byte pieces; zero outside the volume)
    auto issue_tile = [&](int w, int st) {
        int bl, tx, ty, tz;
        tile_coords(w, bl, tx, ty, tz);
        const T* __restrict__ img = ds_dout + (b0 + bl) * grid.cells;
        const uint32_t dst0 = smem_u32(ring + (size_t)st * kTileCells);
#pragma unroll
        for (int r = 0; r < kTileCells / VEC / kThreads; ++r) {
            const int i = r * kThreads + (int)threadIdx.x;
            const int x4 = (i % (TX / VEC)) * VEC, y = (i / (TX / VEC)) % TY, z = i / ((TX / VEC) * TY);
            const int gx = tx * TX + x4, gy = ty * TY + y, gz = tz * TZ + z;
            const bool valid = gx < grid.g[0] && gy < grid.g[1] && gz < grid.g[2];      // rows are multiples of VEC: whole piece
            const T* src = valid ? img + ((int64_t)gz * grid.g[1] + gy) * grid.g[0] + gx : ds_dout;
            cp_async16_zfill(dst0 + (uint32_t)i * 16u, src, valid);
        }
    };

    if (threadIdx.x == 0) {
        s_ticket[0] = (int)atomicAdd(work_counter, 1u);
        s_ticket[1] = (int)atomicAdd(work_counter, 1u);
    }
    __syncthreads();
    int w_cur = s_ticket[0], w_next = s_ticket[1];
    if (w_cur >= n_work) return;
    int st = 0;
    issue_tile(w_cur, 0);
    cp_async_commit();
    {
        int bl, tx, ty, tz;
        tile_coords(w_cur, bl, tx, ty, tz);
        build_ranges(rg[0], cnt, tg, bl, tx, ty, tz);
    }

    T acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = T(0);
    T bg_part = T(0);
    int cur_bl = -1;
    Pose<T, N_IN, 3> pose;
    // block-reduce the per-pose sums of pose b0 + cur_bl and add them to the outputs (all threads call it)
    auto flush_pose = [&]() {
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] = warp_sum(acc[v]);
        bg_part = warp_sum(bg_part);
        __syncthreads();
        if (lane == 0) {
#pragma unroll
            for (int v = 0; v < NV; ++v) red[warp][v] = acc[v];
            red[warp][NV] = bg_part;
        }
        __syncthreads();
        if (threadIdx.x <= NV) {
            const int v = threadIdx.x;
            const int64_t b = b0 + cur_bl;
            T r = T(0);
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w) r += red[w][v];
            if (v == NV) { if (d_background) red_add(d_background + b, r); }
            else if (r != T(0)) {
                if (v < NR) red_add(d_rotation + b * NR + v, r);
                else if (v < NR + 3) red_add(d_translation + b * 3 + (v - NR), r);
                else if (d_out_weight) red_add(d_out_weight + b, r);
            }
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] = T(0);
        bg_part = T(0);
    };

    for (int iter = 0;; ++iter) {
        // the ticket after next: written to a slot that alternates per iteration, read after this iteration's last barrier
        if (threadIdx.x == 0) s_ticket[2 + (iter & 1)] = (int)atomicAdd(work_counter, 1u);
        if (w_next < n_work) issue_tile(w_next, st ^ 1);
        cp_async_commit();
        RangeLoad rl_next;
        rl_next.s[0] = rl_next.s[1] = rl_next.e[0] = rl_next.e[1] = 0u;
        if (w_next < n_work) {              // warp 0 issues the loads now and uses them after this tile's entries
            int bl, tx, ty, tz;
            tile_coords(w_next, bl, tx, ty, tz);
            rl_next = ranges_issue(cnt, tg, bl, tx, ty, tz);
        }
        cp_async_wait<1>();                 // this thread's pieces of the current tile have landed
        __syncthreads();                    // everyone's pieces + rg[st] visible

        int bl, tx, ty, tz;
        tile_coords(w_cur, bl, tx, ty, tz);
        if (bl != cur_bl) {
            if (cur_bl >= 0) flush_pose();
            cur_bl = bl;
            load_pose<T, N_IN, 3>(pose, rotation, translation, out_weight, b0 + bl);
        }
        const T* __restrict__ tile = ring + (size_t)st * kTileCells;
        const int ox = tx * TX, oy = ty * TY, oz = tz * TZ;
        if (d_background) {                 // src/raster_pullback.jl:78 - from the staged tile (zero outside the volume)
#pragma unroll
            for (int r = 0; r < kTileCells / VEC / kThreads; ++r) {
                const Pack pk = reinterpret_cast<const Pack*>(tile)[r * kThreads + threadIdx.x];
#pragma unroll
                for (int k = 0; k < VEC; ++k) bg_part += pk.v[k];
            }
        }
        const Ranges& rgc = rg[st];
        const uint32_t total = rgc.pref[64];
        if (total) {
            EntryCursor cur{&rgc, entries, 0};
            uint32_t e = threadIdx.x;
            uint32_t idx_a = e < total ? cur.fetch(e) : 0u;
            uint32_t idx_b = e + kThreads < total ? cur.fetch(e + kThreads) : 0u;
            Pt4<T> qn = pts4[idx_a];
            for (; e < total; e += kThreads) {
                const Pt4<T> q = qn;
                const uint32_t idx = idx_a;
                idx_a = idx_b;
                if (e + kThreads < total) qn = pts4[idx_a];
                if (e + 2 * kThreads < total) idx_b = cur.fetch(e + 2 * kThreads);
                T x[N_IN];
                load_xyz<T, N_IN>(x, q);
                int i0[3];
                T dl[3];
                if (!stencil<T, N_IN, 3>(x, pose, grid, i0, dl)) continue;
                const int lx = i0[0] - ox, ly = i0[1] - oy, lz = i0[2] - oz;
                // corner values in bit order (x = bit 0).  Corners of other tiles are that tile's CTA's job (everything is
                // linear in G); corners outside the volume read the zero fill = skipped (src/raster_pullback.jl:51)
                T G[8];
                const bool interior = (unsigned)lx < (unsigned)(TX - 1) && (unsigned)ly < (unsigned)(TY - 1) && (unsigned)lz < (unsigned)(TZ - 1);
                const int off = (lz * TY + ly) * TX + lx;
                if (interior) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) G[c] = tile[off + (c & 1) + ((c >> 1) & 1) * TX + ((c >> 2) & 1) * TX * TY];
                } else {
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const int cx = lx + (c & 1), cy = ly + ((c >> 1) & 1), cz = lz + ((c >> 2) & 1);
                        const bool in = (unsigned)cx < (unsigned)TX && (unsigned)cy < (unsigned)TY && (unsigned)cz < (unsigned)TZ;
                        G[c] = in ? tile[(cz * TY + cy) * TX + cx] : T(0);
                    }
                }
                // s = sum_c W_c G_c (src/raster_pullback.jl:55-58) and gk[n] = d s / d dl_n (:60-65, :150-160), factorised:
                // lerp / difference along x, then y, then z
                T ax[4], dx[4];
#pragma unroll
                for (int yz = 0; yz < 4; ++yz) {
                    dx[yz] = G[2 * yz + 1] - G[2 * yz];
                    ax[yz] = fma(dl[0], dx[yz], G[2 * yz]);
                }
                T by[2], dy[2], ex[2];
#pragma unroll
                for (int z = 0; z < 2; ++z) {
                    dy[z] = ax[2 * z + 1] - ax[2 * z];
                    by[z] = fma(dl[1], dy[z], ax[2 * z]);
                    ex[z] = fma(dl[1], dx[2 * z + 1] - dx[2 * z], dx[2 * z]);
                }
                T gk[3];
                gk[2] = by[1] - by[0];
                const T s = fma(dl[2], gk[2], by[0]);
                gk[1] = fma(dl[2], dy[1] - dy[0], dy[0]);
                gk[0] = fma(dl[2], ex[1] - ex[0], ex[0]);
                acc[NV - 1] += s * q.w;                           // d_out_weight,   src/raster_pullback.jl:57
                const T f = pose.ow * q.w;                        // :60
                T scaled[3];
#pragma unroll
                for (int n = 0; n < 3; ++n) {
                    scaled[n] = (f * gk[n]) * grid.scale[n];      // :67
                    acc[NR + n] += scaled[n];                     // d_translation, :68
                }
                T dpt[3] = {T(0), T(0), T(0)};
#pragma unroll
                for (int j = 0; j < N_IN; ++j) {
                    T d = T(0);
#pragma unroll
                    for (int n = 0; n < 3; ++n) {
                        acc[n + j * 3] += scaled[n] * x[j];       // d_rotation, :69
                        d += pose.R[n][j] * scaled[n];            // R' * scaled, :70
                    }
                    dpt[j] = d;
                }
                // pose-sum of d_points (:71, :141) and d_point_weight (:58, :146): one 16-byte reduction into the packed,
                // L2-resident buffer (sorted point order)
                red_add4(acc4 + idx, dpt[0], dpt[1], dpt[2], s * pose.ow);
            }
        }
        ranges_finish(rg[st ^ 1], rl_next);
        __syncthreads();                    // stage st may be refilled; rg[st ^ 1] and the new ticket are visible
        w_cur = w_next;
        w_next = s_ticket[2 + (iter & 1)];
        st ^= 1;
        if (w_cur >= n_work) break;
    }
    flush_pose();
}

// 6. packed, sorted-order gradients -> d_points (N_in, P) and d_point_weight (P) in the caller's point order
template <typename T, int N_IN>
__global__ void __launch_bounds__(256) unpermute_kernel(const Pt4<T>* __restrict__ acc4, const int32_t* __restrict__ perm, int64_t P,
                                                        T* __restrict__ d_points, T* __restrict__ d_point_weight) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += stride) {
        const Pt4<T> a = acc4[i];
        const int64_t p = perm[i];
        d_points[p * N_IN] = a.x;
        if constexpr (N_IN > 1) d_points[p * N_IN + 1] = a.y;
        if constexpr (N_IN > 2) d_points[p * N_IN + 2] = a.z;
        if (d_point_weight) d_point_weight[p] = a.w;
    }
}

// ---------------------------------------------------------------------------------------------------------
// host side: workspace plan + the shared binning driver
// ---------------------------------------------------------------------------------------------------------
struct Plan {
    bool ok = false;
    int sort_bits = 0;
    TileGeom tg{};
    int64_t group = 0;             // poses binned per pass
    ScanRegion sort_scan, tile_scan;
    size_t off_keys = 0, off_perm = 0, off_pts4 = 0, off_acc4 = 0, off_entries = 0, total = 0;
};

inline Plan make_plan(int n_in, const int64_t* grid, int64_t P, int64_t B, int sizeof_T, bool pullback) {
    Plan pl;
    if (n_in > 3 || P < 1 || B < 1 || P >= ((int64_t)1 << 30)) return pl;
    const int TS[3] = {TX, TY, TZ};
    int64_t n_tiles = 1;
    for (int k = 0; k < 3; ++k) {
        if (grid[k] >= ((int64_t)1 << 20)) return pl;
        pl.tg.nt[k] = (int)((grid[k] + TS[k] - 1) / TS[k]);
        n_tiles *= pl.tg.nt[k];
    }
    if (n_tiles * 8 > ((int64_t)1 << 24)) return pl;
    pl.tg.n_tiles = (int)n_tiles;
    // poses per pass: at most 2^27 entries (512 MB), 2^24 counters and 65535 (gridDim.y) poses
    int64_t group = B;
    if (group > ((int64_t)1 << 27) / P) group = ((int64_t)1 << 27) / P;
    if (group > ((int64_t)1 << 24) / (n_tiles * 8)) group = ((int64_t)1 << 24) / (n_tiles * 8);
    if (group > 65535) group = 65535;
    if (group < 1) group = 1;
    if (group * n_tiles > (int64_t)0x7fffffff) return pl;
    pl.group = group;
    // pre-sort cells: ~4 points per cell, 2..6 bits per dimension (measured on B200, 1 M points: 2^15 cells make the
    // histogram and scatter atomics three times slower than 2^18 cells - same-address contention at L2)
    int bits = 2;
    while (bits < 6 && ((int64_t)4 << (n_in * (bits + 1))) <= P) ++bits;
    pl.sort_bits = bits;
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    size_t o = 256;
    pl.off_keys = o;    o = al(o + sizeof(uint32_t) * (size_t)P);
    pl.off_perm = o;    o = al(o + sizeof(int32_t) * (size_t)P);
    pl.off_pts4 = o;    o = al(o + (size_t)sizeof_T * 4 * (size_t)P);
    pl.off_acc4 = o;    o = al(o + (pullback ? (size_t)sizeof_T * 4 * (size_t)P : 0));
    pl.sort_scan = make_scan_region(o, (int64_t)1 << (bits * n_in));
    o = pl.sort_scan.off_ticket + pl.sort_scan.bytes;
    pl.tile_scan = make_scan_region(o, group * n_tiles * 8);
    o = pl.tile_scan.off_ticket + pl.tile_scan.bytes;
    pl.off_entries = o; o = al(o + sizeof(uint32_t) * (size_t)(P * group));
    pl.total = o;
    pl.ok = true;
    return pl;
}

template <typename T, int N_IN>
static int presort(const T* points, const T* point_weight, int64_t P, char* ws, const Plan& pl, bool zero_acc,
                   const DeviceInfo& dev, cudaStream_t stream) {
    uint32_t* keys = reinterpret_cast<uint32_t*>(ws + pl.off_keys);
    uint32_t* counts = reinterpret_cast<uint32_t*>(ws + pl.sort_scan.off_data);
    int rc = clear_scan_region(ws, pl.sort_scan, stream);
    if (rc != DPR_OK) return rc;
    int64_t blocks = (P + 255) / 256;
    if (blocks > (int64_t)dev.sm_count * 16) blocks = (int64_t)dev.sm_count * 16;
    {
        LaunchScope scope("tile3_sort_count", stream);
        bin_count_kernel<T, N_IN><<<(unsigned)blocks, 256, 0, stream>>>(points, P, pl.sort_bits, keys, counts);
    }
    rc = launch_scan(ws, pl.sort_scan, stream, "tile3_sort_scan");
    if (rc != DPR_OK) return rc;
    {
        LaunchScope scope("tile3_sort_scatter", stream);
        sort_scatter4_kernel<T, N_IN><<<(unsigned)blocks, 256, 0, stream>>>(
            points, point_weight, P, keys, counts, reinterpret_cast<int32_t*>(ws + pl.off_perm),
            reinterpret_cast<Pt4<T>*>(ws + pl.off_pts4), zero_acc ? reinterpret_cast<Pt4<T>*>(ws + pl.off_acc4) : nullptr,
            reinterpret_cast<uint32_t*>(ws + pl.sort_scan.off_ticket));
    }
    DPR_CUDA_TRY(cudaGetLastError());
    return DPR_OK;
}

// count + scan + scatter for poses [b0, b0 + nb)
template <typename T, int N_IN>
static int bin_poses(const T* rotation, const T* translation, const Grid<T, 3>& grid, int64_t P, int64_t b0, int64_t nb,
                     char* ws, const Plan& pl, cudaStream_t stream) {
    ScanRegion sr = pl.tile_scan;
    sr.n = nb * pl.tg.n_tiles * 8;
    sr.chunks = (int)((sr.n + kScanChunk - 1) / kScanChunk);
    int rc = clear_scan_region(ws, sr, stream);
    if (rc != DPR_OK) return rc;
    const Pt4<T>* pts4 = reinterpret_cast<const Pt4<T>*>(ws + pl.off_pts4);
    uint32_t* cnt = reinterpret_cast<uint32_t*>(ws + sr.off_data);
    uint32_t* entries = reinterpret_cast<uint32_t*>(ws + pl.off_entries);
    const dim3 gridDim3((unsigned)((P + 1023) / 1024), (unsigned)nb);
    {
        LaunchScope scope("tile3_bin_count", stream);
        tile_bin_kernel<T, N_IN, false><<<gridDim3, 256, 0, stream>>>(pts4, (int)P, rotation, translation, grid, pl.tg, cnt, entries, b0);
    }
    rc = launch_scan(ws, sr, stream, "tile3_bin_scan");
    if (rc != DPR_OK) return rc;
    {
        LaunchScope scope("tile3_bin_scatter", stream);
        tile_bin_kernel<T, N_IN, true><<<gridDim3, 256, 0, stream>>>(pts4, (int)P, rotation, translation, grid, pl.tg, cnt, entries, b0);
    }
    DPR_CUDA_TRY(cudaGetLastError());
    return DPR_OK;
}

template <typename T>
static Grid<T, 3> make_grid3(const int64_t* g) {
    Grid<T, 3> grid;
    grid.cells = 1;
    for (int k = 0; k < 3; ++k) {
        grid.g[k] = (int)g[k];
        grid.scale[k] = T(g[k]) / T(2);  // src/raster.jl:25, src/raster_pullback.jl:29
        grid.cells *= g[k];
    }
    return grid;
}

// Is the tile path worth its launches?  It removes the init pass over `out` / the separate d_background pass and the
// L2 round trips of the corner updates, which pays when a good part of the cells is touched; sparse clouds in huge
// volumes (README row 5: 1e5 points in 1024^3) are bound by the one compulsory pass either way and keep the
// point-parallel kernels.
inline bool worthwhile(const int64_t* grid, int64_t P, int64_t B) {
    const double cells = (double)grid[0] * (double)grid[1] * (double)grid[2];
    return P >= 16384 && (double)P * 8.0 >= 0.02 * cells && cells * (double)B >= (double)(1 << 22);
}

}  // namespace t3
}  // namespace dpr
