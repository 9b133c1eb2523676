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

#include <cstddef>
#include <type_traits>

#include "dpr_common.cuh"
#include "dpr_internal.h"
#include "dpr_sort.cuh"

namespace dpr {
namespace t3 {

#ifndef DPR_T3_TY
#define DPR_T3_TY 32
#endif
#ifndef DPR_T3_TZ
#define DPR_T3_TZ 16
#endif
#ifndef DPR_T3_THREADS
#define DPR_T3_THREADS 256
#endif
#ifndef DPR_T3_BATCH
#define DPR_T3_BATCH 4
#endif
#ifndef DPR_T3_MAXCTAS
#define DPR_T3_MAXCTAS 4
#endif
constexpr int TX = 32, TY = DPR_T3_TY, TZ = DPR_T3_TZ;   // tile extent in cells (x contiguous: one 128-byte line per Float32 row)
constexpr int kTileCells = TX * TY * TZ;
constexpr int kThreads = DPR_T3_THREADS;
constexpr int kBatch = DPR_T3_BATCH;              // entries (and then points) a thread has in flight at once
constexpr uint32_t kNoKey = 0xffffffffu;

template <typename T>
struct alignas(4 * sizeof(T)) Pt4 { T x, y, z, w; };
constexpr int kStatMaxBits = 4, kStatBad = 5, kStatSum = 6;     // see "helpers shared by the tile kernels"

// ---------------------------------------------------------------------------------------------------------
// Binning cache (DPR_OPT_BINNING_CACHE).  The pre-sort and the per-pose bins depend on points, point_weight, rotation and
// translation only - exactly what a forward call and the pullback that follows it share (the rrule calls raster, then
// raster_pullback! with the same arguments, ext/DiffPointRasterisationChainRulesCoreExt.jl:56-61).  With the option on, the
// first 256 bytes of the workspace hold a header with a 128-bit hash of those inputs; every call hashes its inputs on
// the device (one pass over ~P (N_in + 1) values), and when hash, shapes and the "complete" mark agree, all binning
// kernels of the call return immediately (the decision is taken on the device: nothing synchronises).  Contract for the
// caller: zero the first 256 bytes of a workspace when it is allocated, and leave the workspace alone between calls.
// Every library path that writes into a workspace without going through the cache clears the mark first.
// ---------------------------------------------------------------------------------------------------------
struct CacheHeader {
    unsigned long long magic, hash[2], params, valid;      // persistent between calls
    unsigned long long acc[2];                              // this call's hash accumulators (zeroed by the host)
    unsigned int done_blocks, skip;                         // scratch; skip = 1: the cached bins are valid for this call
};
constexpr unsigned long long kCacheMagic = 0x4450523354494c45ull;   // "DPR3TILE"

template <typename T>
__device__ __forceinline__ unsigned long long raw_bits(T v) {
    if constexpr (sizeof(T) == 4) return (unsigned long long)__float_as_uint((float)v);
    else return (unsigned long long)__double_as_longlong((double)v);
}
template <typename T>
__global__ void __launch_bounds__(256) hash_inputs_kernel(const T* __restrict__ points, int64_t n_pts, const T* __restrict__ pw, int64_t n_pw,
                                                          const T* __restrict__ rot, int64_t n_rot, const T* __restrict__ tr, int64_t n_tr,
                                                          unsigned long long params, CacheHeader* __restrict__ h) {
    unsigned long long h1 = 0, h2 = 0;
    // position-dependent terms, summed: independent of the order of the threads, dependent on the order of the data.  Two
    // multiply-xorshift rounds per element (the full splitmix64 finaliser on both words made this pass compute bound).
    auto term = [&](T v, int64_t i) {
        unsigned long long t = (raw_bits(v) ^ ((unsigned long long)(i + 1) * 0x9e3779b97f4a7c15ull)) * 0xbf58476d1ce4e5b9ull;
        t ^= t >> 29;
        h1 += t;
        t *= 0x94d049bb133111ebull;
        h2 += t ^ (t >> 32);
    };
    constexpr int VEC = 16 / sizeof(T);
    struct alignas(16) Pack { T v[VEC]; };
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    // one array: elements [0, n) hash at positions base + i.  16-byte loads, two in flight, where the array is aligned
    // (the value of the hash does not depend on which path read an element)
    auto hash_array = [&](const T* __restrict__ a, int64_t n, int64_t base) {
        if (n <= 0) return;
        int64_t n_vec = 0;
        if ((reinterpret_cast<uintptr_t>(a) & 15u) == 0) {
            n_vec = n / VEC;
            const Pack* __restrict__ av = reinterpret_cast<const Pack*>(a);
            for (int64_t g = tid; g < n_vec; g += 2 * stride) {
                const Pack p0 = av[g];
                Pack p1 = p0;
                const bool two = g + stride < n_vec;
                if (two) p1 = av[g + stride];
#pragma unroll
                for (int k = 0; k < VEC; ++k) term(p0.v[k], base + g * VEC + k);
                if (two) {
#pragma unroll
                    for (int k = 0; k < VEC; ++k) term(p1.v[k], base + (g + stride) * VEC + k);
                }
            }
        }
        for (int64_t i = n_vec * VEC + tid; i < n; i += stride) term(__ldg(a + i), base + i);
    };
    hash_array(points, n_pts, 0);
    hash_array(pw, n_pw, n_pts);
    hash_array(rot, n_rot, n_pts + n_pw);
    hash_array(tr, n_tr, n_pts + n_pw + n_rot);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        h1 += __shfl_xor_sync(0xffffffffu, h1, o);
        h2 += __shfl_xor_sync(0xffffffffu, h2, o);
    }
    __shared__ unsigned long long s1[8], s2[8];
    __shared__ bool last;
    if ((threadIdx.x & 31) == 0) { s1[threadIdx.x >> 5] = h1; s2[threadIdx.x >> 5] = h2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { h1 += s1[w]; h2 += s2[w]; }
        atomicAdd(&h->acc[0], h1);
        atomicAdd(&h->acc[1], h2);
        __threadfence();
        last = atomicAdd(&h->done_blocks, 1u) == gridDim.x - 1;
        if (last) {
            __threadfence();
            const unsigned long long a0 = *reinterpret_cast<volatile unsigned long long*>(&h->acc[0]);
            const unsigned long long a1 = *reinterpret_cast<volatile unsigned long long*>(&h->acc[1]);
            const bool hit = h->magic == kCacheMagic && h->valid == 1ull && h->params == params && h->hash[0] == a0 && h->hash[1] == a1;
            if (!hit) {
                h->magic = kCacheMagic;
                h->hash[0] = a0;
                h->hash[1] = a1;
                h->params = params;
                h->valid = 0ull;        // set by the tile kernel that runs after the binning kernels of this call
            }
            h->skip = hit ? 1u : 0u;
        }
    }
}
// zeroes [ptr, ptr + n16 * 16) unless the cache hit
static __global__ void __launch_bounds__(256) clear_unless_cached_kernel(uint4* __restrict__ ptr, int64_t n16, const unsigned int* __restrict__ skip) {
    if (skip && *skip) return;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) ptr[i] = make_uint4(0u, 0u, 0u, 0u);
}

// (the single-pass scan - scan_lookback_kernel, ScanRegion, launch_scan - lives in dpr_sort.cuh)

// ---------------------------------------------------------------------------------------------------------
// 1. pre-sort: keys + histogram come from bin_count_kernel (dpr_sort.cuh); this scatter writes the packed copy
// ---------------------------------------------------------------------------------------------------------
template <typename T, int N_IN>
__global__ void __launch_bounds__(256) sort_count_kernel(const T* __restrict__ points, int64_t P, int bits, uint32_t* __restrict__ keys,
                                                         uint32_t* __restrict__ counts, const unsigned int* __restrict__ skip) {
    if (skip && *skip) return;                 // binning cache hit (uniform)
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += stride) {
        const uint32_t k = morton_key<T, N_IN>(points, p, bits);
        keys[p] = k;
        atomicAdd(counts + k, 1u);
    }
}
template <typename T, int N_IN>
__global__ void __launch_bounds__(256) sort_scatter4_kernel(const T* __restrict__ points, const T* __restrict__ point_weight,
                                                            int64_t P, const uint32_t* __restrict__ keys,
                                                            uint32_t* __restrict__ offsets, int32_t* __restrict__ perm,
                                                            Pt4<T>* __restrict__ pts4, uint32_t* __restrict__ stats,
                                                            const unsigned int* __restrict__ skip) {
    if (skip && *skip) return;                 // binning cache hit (uniform)
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
        perm[p] = (int32_t)pos;         // INVERSE permutation (coalesced store): original index -> sorted position
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
    unsigned home[3];
    unsigned pat = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        // tile of the upper corner (i0 + 1 >= 0) and of the lower one; unsigned divisions by a power of two are shifts.  A lower
        // corner outside the grid (i0 = -1) makes the upper corner's tile the home, an upper corner outside the grid (or in
        // the same tile) leaves the pattern bit clear.
        const unsigned t_hi = (unsigned)(i0[k] + 1) / (unsigned)TS[k];
        const unsigned t_lo = i0[k] >= 0 ? (unsigned)i0[k] / (unsigned)TS[k] : t_hi;
        home[k] = t_lo;
        if (i0[k] + 1 < g[k] && t_hi != t_lo) pat |= 1u << k;
    }
    const uint32_t tile = (home[2] * (uint32_t)tg.nt[1] + home[1]) * (uint32_t)tg.nt[0] + home[0];
    return (((uint32_t)pose_local * (uint32_t)tg.n_tiles + tile) << 3) | pat;
}

// pass 1: key of every (point, pose) pair - kept for pass 2 - and the histogram
template <typename T, int N_IN>
__global__ void __launch_bounds__(256) tile_count_kernel(const Pt4<T>* __restrict__ pts4, int P, const T* __restrict__ rotation,
                                                         const T* __restrict__ translation, Grid<T, 3> grid, TileGeom tg,
                                                         uint32_t* __restrict__ cnt, uint32_t* __restrict__ keys, int64_t b0,
                                                         const unsigned int* __restrict__ skip) {
    if (skip && *skip) return;                 // binning cache hit (uniform)
    constexpr int K = 4;
    const int bl = blockIdx.y;
    Pose<T, N_IN, 3> pose;
    load_pose<T, N_IN, 3>(pose, rotation, translation, nullptr, b0 + bl);
    const int lane = threadIdx.x & 31;
    uint32_t* __restrict__ my_keys = keys + (size_t)bl * (size_t)P;
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
            my_keys[p] = key;
        }
        // spatially sorted points: the 32 lanes of a warp share a handful of keys -> one atomic per distinct key
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        if (key != kNoKey && lane == __ffs(peers) - 1) atomicAdd(cnt + key, (uint32_t)__popc(peers));
    }
}
constexpr int kScatterK = 8;
// pass 2 (after the scan): the sorted point index of every pair goes to its list; no transform, the keys are re-read
static __global__ void __launch_bounds__(256) tile_scatter_kernel(int P, uint32_t* __restrict__ cnt, const uint32_t* __restrict__ keys,
                                                                  uint32_t* __restrict__ entries, const unsigned int* __restrict__ skip) {
    if (skip && *skip) return;                 // binning cache hit (uniform)
    constexpr int K = kScatterK;               // the pass is two dependent round trips per CTA: long CTAs, few waves
    const uint32_t* __restrict__ my_keys = keys + (size_t)blockIdx.y * (size_t)P;
    const int lane = threadIdx.x & 31;
    uint32_t key[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int p = (blockIdx.x * K + k) * 256 + (int)threadIdx.x;
        key[k] = p < P ? __ldg(my_keys + p) : kNoKey;
    }
    // all K atomics of a thread are in flight before the first result is needed (one L2 round trip, not K)
    unsigned peers[K];
    uint32_t base[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        peers[k] = __match_any_sync(0xffffffffu, key[k]);
        base[k] = 0;
        if (key[k] != kNoKey && lane == __ffs(peers[k]) - 1) base[k] = atomicAdd(cnt + key[k], (uint32_t)__popc(peers[k]));
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int p = (blockIdx.x * K + k) * 256 + (int)threadIdx.x;
        const uint32_t b = __shfl_sync(0xffffffffu, base[k], __ffs(peers[k]) - 1);
        if (key[k] != kNoKey) entries[b + __popc(peers[k] & ((1u << lane) - 1u))] = (uint32_t)p;
    }
}

// ---------------------------------------------------------------------------------------------------------
// 5. the tile kernels: one CTA per (pose, tile) work item, 256 threads, 32 x 32 x 16 cells = 64 KB of Float32 per tile, so
// that three CTAs share an SM (78 / 77 registers) and the hardware scheduler overlaps their phases (sub-list table, entry
// and point gathers, accumulation, flush).
//
// What was measured on config 3 on the way here (all on the same B200 pool):
//   v2  CTA per item, float CAS atomics, div/mod in the flush        fwd 467 us (273 M warp instructions, issue 51 %)
//   v2b + fixed-point atomics                                          fwd 414 us (324 M, issue 68 %, 39 warps / SM)
//   v3  persistent CTAs, warp 0 does the bookkeeping between barriers  fwd 421 us, pullback 498 us (barrier stalls 32 %)
//   v5  warp-specialised persistent CTAs (metadata producer warp, TMA  fwd 535 us, pullback 464 us: 190 M / 150 M
//       producer warp, mbarrier rings, cross-item look-ahead)          instructions but 72 - 93 registers -> 20 - 27 warps
//                                                                      per SM, issue 39 %; forcing 56 - 64 registers
//                                                                      (spills) made it 675 / 889 us
//   v6  (this file) one CTA per item with the lean pieces              fwd 341 us, pullback 308 us
// The instruction-lean pieces of v3 - v5 (division-free cell map, compact sub-list table, fixed-point accumulation,
// factorised trilinear gradient, tensor-map TMA tile loads) are kept; the bookkeeping went back to the hardware.
// With no points at all the kernels run at the HBM floor of their one pass over the volume (163 / 177 us for 1.07 GB);
// every million points (16.8 M entries) adds 155 / 161 us that do not overlap with that pass: the entries are bound by
// issue slots and the shared-memory pipe, not by load latency (deeper prefetching changes nothing), and three CTAs per SM,
// each in one phase at a time, overlap the phases only statistically (tools/exp_tile3d_scaling.py).
// The DPR_T3_* macros exist for tools/build_variants.py (experiment builds with other tile shapes); the defaults are the
// product.
// ---------------------------------------------------------------------------------------------------------
constexpr int kMaxSlots = 27;

// the 27 (lower-neighbour offset d, pattern) combinations with pattern containing d; combination 0 is the tile's own
// pattern-0 list (stencils entirely inside the tile)
// (3 bits per lane, packed into immediates: a lane-indexed __constant__ load is serialised per distinct address)
constexpr unsigned char kComboD[27] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 4, 4, 4, 4, 5, 5, 6, 6, 7};
constexpr unsigned char kComboPat[27] = {0, 1, 2, 3, 4, 5, 6, 7, 1, 3, 5, 7, 2, 3, 6, 7, 3, 7, 4, 5, 6, 7, 5, 7, 6, 7, 7};
__host__ __device__ constexpr unsigned long long pack3(const unsigned char (&v)[27], int first, int last) {
    unsigned long long r = 0;
    for (int i = first; i < last; ++i) r |= (unsigned long long)v[i] << (3 * (i - first));
    return r;
}
constexpr unsigned long long kComboDLo = pack3(kComboD, 0, 21), kComboDHi = pack3(kComboD, 21, 27);
constexpr unsigned long long kComboPatLo = pack3(kComboPat, 0, 21), kComboPatHi = pack3(kComboPat, 21, 27);
__device__ __forceinline__ int combo_lookup(unsigned long long lo, unsigned long long hi, int lane) {
    return (int)(((lane < 21 ? lo >> (3 * lane) : hi >> (3 * (lane - 21)))) & 7ull);      // lanes >= 27 read zeros
}

// compact table of the non-empty sub-lists of one item, built by warp 0
struct SlotTable {
    int n_slots;
    uint32_t total;                // entries over all sub-lists
    uint32_t n_interior;           // length of the tile's own pattern-0 list (always slot 0 when non-empty)
    uint32_t base[kMaxSlots];      // sub-list s holds concatenated entries e in [end[s-1], end[s]) at entries[base[s] + e]
    uint32_t end[kMaxSlots];
};
__device__ __forceinline__ void build_slot_table(SlotTable& tab, const uint32_t* __restrict__ cnt, const TileGeom& tg, int bl,
                                                 int tx, int ty, int tz) {
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    constexpr unsigned long long d_lo = kComboDLo, d_hi = kComboDHi, p_lo = kComboPatLo, p_hi = kComboPatHi;
    const int d = combo_lookup(d_lo, d_hi, lane), pat = combo_lookup(p_lo, p_hi, lane);
    const int nx = tx - (d & 1), ny = ty - ((d >> 1) & 1), nz = tz - ((d >> 2) & 1);
    uint32_t s = 0, e = 0;
    if (lane < kMaxSlots && nx >= 0 && ny >= 0 && nz >= 0) {
        const uint32_t tile = (uint32_t)((nz * tg.nt[1] + ny) * tg.nt[0] + nx);
        const uint32_t key = (((uint32_t)bl * (uint32_t)tg.n_tiles + tile) << 3) | (uint32_t)pat;
        e = __ldg(cnt + key);                           // after the scatter pass cnt[key] is the END of list `key`
        s = key ? __ldg(cnt + key - 1) : 0u;
    }
    const uint32_t len = e - s;
    uint32_t incl = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const unsigned mask = __ballot_sync(0xffffffffu, len != 0u);
    if (len) {
        const int pos = __popc(mask & ((1u << lane) - 1u));
        tab.base[pos] = s - (incl - len);
        tab.end[pos] = incl;
    }
    if (lane == 0) { tab.n_slots = __popc(mask); tab.n_interior = len; }
    if (lane == 31) tab.total = incl;
}

// Walks the compact sub-list table as one sequence: index e of the concatenation -> sorted point index.
struct EntryCursor {
    const SlotTable* t;
    const uint32_t* entries;
    int slot;
    uint32_t bound, base;
    __device__ __forceinline__ void init(const SlotTable* tab, const uint32_t* en) {
        t = tab; entries = en; slot = 0;
        bound = tab->end[0];
        base = tab->base[0];
    }
    __device__ __forceinline__ uint32_t fetch(uint32_t e) {
        if (e >= bound) {
            do { ++slot; bound = t->end[slot]; } while (e >= bound);
            base = t->base[slot];
        }
        return __ldg(entries + (uint32_t)(base + e));       // 32-bit sum: base may have wrapped below zero
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

// Thread <-> tile cell mapping of the flush / tile sums: 16-byte piece r (r = 0 .. kPieces-1) of thread tid is the piece with
// linear index tid + r * kThreads of the x-contiguous tile, i.e. cell (x4, y0 + dy(r), z0 + dz(r)) with x4, y0, z0 fixed per
// thread and dy, dz compile-time functions of r - no division in the loops.
template <typename T>
struct CellMap {
    static constexpr int VEC = 16 / sizeof(T);
    static constexpr int PX = TX / VEC;                            // pieces per row
    static constexpr int kPlanePieces = PX * TY;
    static constexpr int kPieces = kTileCells / VEC / kThreads;    // per thread
    static constexpr bool kMultiY = kPlanePieces > kThreads;       // a plane has more pieces than the CTA has threads
    static constexpr int RY = kMultiY ? kPlanePieces / kThreads : 1;          // pieces per plane and thread
    static constexpr int YS = kThreads / PX;                                   // their distance in y
    static constexpr int kZStep = kMultiY ? 1 : kThreads / kPlanePieces;
    static_assert(kThreads % PX == 0 && (kMultiY ? kPlanePieces % kThreads == 0 : kThreads % kPlanePieces == 0) &&
                  kPieces * kThreads * VEC == kTileCells, "tile / CTA shape mismatch");
    static constexpr __host__ __device__ int dy(int r) { return kMultiY ? (r % RY) * YS : 0; }
    static constexpr __host__ __device__ int dz(int r) { return kMultiY ? r / RY : r * kZStep; }
    int x4, y, z0;
    __device__ __forceinline__ CellMap() {
        const int tid = threadIdx.x;
        x4 = (tid % PX) * VEC;
        y = kMultiY ? tid / PX : (tid / PX) % TY;
        z0 = kMultiY ? 0 : tid / kPlanePieces;
    }
};

// ---------------------------------------------------------------------------------------------------------
// 5a. forward: accumulate in shared memory, one store per output cell (src/raster.jl:27,36-66)
//
// Float32 accumulates in FIXED POINT on the native 32-bit shared-memory reduction (ATOMS.ADD without return: the float
// atomicAdd is an LDS + ATOMS.CAST.SPIN loop on sm_100a and made the first version issue 470 instructions per warp-entry).
// Same scheme as the 2-d kernels (dpr_forward_fast.cuh): a contribution w * out_weight * point_weight = w * pw' * A
// (pw' = point_weight * 2^-em in [0, 1), A = out_weight * 2^em) is added as the integer rint(w * pw' * Q),
// 2^(F-1) < Q <= 2^F <= 2^22, produced without a conversion (a product with the subnormal whose bit pattern is Q is
// rounded to exactly that integer); the flush multiplies by A / Q.  A tile with fewer than 2^(32-F) entries cannot wrap a
// 32-bit cell; heavier tiles carry a mass checksum that detects a wrapped cell exactly, and the tile is then redone with
// float atomics, as it is for poses / weights that are not eligible (non-positive out_weight, negative or non-finite
// point weights, dynamic range above 64).
// The tile rows are padded by one 16-byte piece and the planes by two: spatially sorted entries put the 32 lanes of a warp
// into a blob a few cells wide, and with a dense 32-float pitch the bank would depend on x only (6 wavefronts per atomic).
// ---------------------------------------------------------------------------------------------------------
template <typename T>
struct FwdTile {
    static constexpr int VEC = 16 / sizeof(T);
    static constexpr int PITCH = TX + VEC;                  // elements per row
    static constexpr int PLANE = TY * PITCH + 2 * VEC;      // elements per z-plane
    static constexpr int SIZE = TZ * PLANE;                 // elements
};

// CTAs per SM the shared-memory tile allows (227 KB per SM, 1 KB reserved per CTA), at most 5: the register budget follows
constexpr int resident_ctas(size_t tile_bytes) { return (int)((227 * 1024) / (tile_bytes + 2048)) < DPR_T3_MAXCTAS ? (int)((227 * 1024) / (tile_bytes + 2048)) : DPR_T3_MAXCTAS; }

template <typename T, int N_IN>
__global__ void __launch_bounds__(kThreads, resident_ctas(sizeof(T) * FwdTile<T>::SIZE))
fwd_tile3d_kernel(const Pt4<T>* __restrict__ pts4, const uint32_t* __restrict__ entries, const uint32_t* __restrict__ cnt,
                  const T* __restrict__ rotation, const T* __restrict__ translation, const T* __restrict__ background,
                  const T* __restrict__ out_weight, T* __restrict__ out, Grid<T, 3> grid, TileGeom tg, int64_t b0,
                  const uint32_t* __restrict__ pw_stats, int64_t P, int fixed_bits, unsigned long long* __restrict__ cache_valid) {
    using CM = CellMap<T>;
    using FT = FwdTile<T>;
    // this kernel runs after the binning kernels of the call: the cached bins are complete (see CacheHeader)
    if (threadIdx.x == 0 && cache_valid && (blockIdx.x | blockIdx.y | blockIdx.z) == 0) *cache_valid = 1ull;
    constexpr int VEC = CM::VEC;
    struct alignas(16) Pack { T v[VEC]; };
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* tile = reinterpret_cast<T*>(smem_raw);
    __shared__ SlotTable tab;
    __shared__ long long scratch[kThreads / 32];
    // launch grid (nt0, nt1, nt2 * poses): one division instead of four
    const int tx = blockIdx.x, ty = blockIdx.y;
    const int bl = blockIdx.z / tg.nt[2], tz = blockIdx.z - bl * tg.nt[2];
    const int64_t b = b0 + bl;
    build_slot_table(tab, cnt, tg, bl, tx, ty, tz);

    const CM cm;
    const int ox = tx * TX, oy = ty * TY, oz = tz * TZ;
    const T bg = background ? __ldg(background + b) : T(0);
    const bool vec_ok = (grid.g[0] % VEC) == 0 && (reinterpret_cast<uintptr_t>(out) % 16) == 0;
    const int64_t plane = (int64_t)grid.g[0] * grid.g[1];
    T* const my_cells = tile + (cm.z0 * FT::PLANE + cm.y * FT::PITCH + cm.x4);      // + r * kZStep * PLANE
    T* __restrict__ dst0 = out + b * grid.cells + ((int64_t)(oz + cm.z0) * grid.g[1] + (oy + cm.y)) * grid.g[0] + (ox + cm.x4);
    const bool col_ok = ox + cm.x4 < grid.g[0];
    const bool full_vec = vec_ok && ox + cm.x4 + VEC <= grid.g[0];
    const int y_lim = grid.g[1] - oy - cm.y, z_lim = grid.g[2] - oz - cm.z0;    // piece r is inside the volume iff dy(r) < y_lim, dz(r) < z_lim
    float inv_q = 0.f;

    // stores value(cell) + bg for this thread's cells inside the volume.
    // MODE 0: background only (no shared-memory access); 1: float cells; 2: fixed-point cells; 3: fixed point + cell sum,
    // and the cells are zeroed for a possible second pass
    auto flush = [&](auto mode_tag) -> long long {
        constexpr int MODE = decltype(mode_tag)::value;
        unsigned long long cells = 0;
#pragma unroll
        for (int r = 0; r < CM::kPieces; ++r) {
            T* dst = dst0 + (CM::dz(r) * plane + (int64_t)CM::dy(r) * grid.g[0]);
            Pack pk;
            if constexpr (MODE == 0) {
#pragma unroll
                for (int k = 0; k < VEC; ++k) pk.v[k] = bg;
            } else {
                Pack* cell = reinterpret_cast<Pack*>(my_cells + CM::dz(r) * FT::PLANE + CM::dy(r) * FT::PITCH);
                pk = *cell;
                if constexpr (MODE >= 2 && sizeof(T) == 4) {
                    if constexpr (MODE == 3) {
                        Pack z;
#pragma unroll
                        for (int k = 0; k < VEC; ++k) z.v[k] = T(0);
                        *cell = z;
                    }
#pragma unroll
                    for (int k = 0; k < VEC; ++k) {
                        const uint32_t u = __float_as_uint(pk.v[k]);
                        if constexpr (MODE == 3) cells += u;
                        pk.v[k] = fmaf((float)u, inv_q, bg);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < VEC; ++k) pk.v[k] += bg;
                }
            }
            if (!col_ok || CM::dy(r) >= y_lim || CM::dz(r) >= z_lim) continue;
            if (full_vec) {
                __stcs(reinterpret_cast<float4*>(dst), *reinterpret_cast<const float4*>(&pk));
            } else {
#pragma unroll
                for (int k = 0; k < VEC; ++k) if (ox + cm.x4 + k < grid.g[0]) dst[k] = pk.v[k];
            }
        }
        return (long long)cells;
    };
    auto zero_tile = [&]() {
#pragma unroll
        for (int r = 0; r < CM::kPieces; ++r) {
            Pack z;
#pragma unroll
            for (int k = 0; k < VEC; ++k) z.v[k] = T(0);
            *reinterpret_cast<Pack*>(my_cells + CM::dz(r) * FT::PLANE + CM::dy(r) * FT::PITCH) = z;
        }
    };
    zero_tile();         // (also the padding stays untouched: it is never read)
    __syncthreads();
    const uint32_t total = tab.total;
    if (total == 0) {            // nothing lands here: the tile is the background (src/raster.jl:27)
        flush(std::integral_constant<int, 0>{});
        return;
    }
    Pose<T, N_IN, 3> pose;
    load_pose<T, N_IN, 3>(pose, rotation, translation, out_weight, b);

    // ---- fixed-point eligibility and scale (uniform over the CTA) ---------------------------------------------
    bool fixed = false;
    float wq_den = 0.f, pw_scale = 1.f;
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

    long long mass = 0;
    const uint32_t n_int = tab.n_interior;
    // one pass over the item's entries.  FIXED: accumulation mode.  Entries below n_int are the tile's own pattern-0
    // list, whose stencils lie inside the tile (only the volume's faces can still clip them: the checked path)
    auto accumulate = [&](auto fixed_tag) {
        constexpr bool FIXED = decltype(fixed_tag)::value;
        if (threadIdx.x >= total) return;
        EntryCursor cur;
        cur.init(&tab, entries);
        const uint32_t tile_s = smem_u32(tile);
        // kBatch entries, then their kBatch points, are in flight at once: two dependent round trips per batch instead of
        // two per entry (an item has about four entries per thread, so a rolling prefetch never reaches steady state)
        uint32_t idx[kBatch], idx_next[kBatch];
#pragma unroll
        for (int j = 0; j < kBatch; ++j) idx_next[j] = threadIdx.x + j * kThreads < total ? cur.fetch(threadIdx.x + j * kThreads) : 0u;
        for (uint32_t e0 = threadIdx.x; e0 < total; e0 += kBatch * kThreads) {
#pragma unroll
            for (int j = 0; j < kBatch; ++j) idx[j] = idx_next[j];
            Pt4<T> qb[kBatch];
#pragma unroll
            for (int j = 0; j < kBatch; ++j) qb[j] = pts4[idx[j]];
            if (e0 + kBatch * kThreads < total) {       // the next batch's entries travel while this batch is computed
#pragma unroll
                for (int j = 0; j < kBatch; ++j)
                    idx_next[j] = e0 + (kBatch + j) * kThreads < total ? cur.fetch(e0 + (kBatch + j) * kThreads) : 0u;
            }
#pragma unroll
            for (int j = 0; j < kBatch; ++j) {
            const uint32_t e = e0 + j * kThreads;
            if (e >= total) break;
            const Pt4<T> q = qb[j];
            T x[N_IN];
            load_xyz<T, N_IN>(x, q);
            int i0[3];
            T dl[3], du[3];
            if (!stencil<T, N_IN, 3>(x, pose, grid, i0, dl)) continue;      // cannot happen: binned with the same arithmetic
#pragma unroll
            for (int k = 0; k < 3; ++k) du[k] = T(1) - dl[k];
            const int lx = i0[0] - ox, ly = i0[1] - oy, lz = i0[2] - oz;
            // corner values in bit order (x = bit 0): w_c * out_weight * point_weight (src/raster.jl:51,63,104-106),
            // or its fixed-point image
            T v[8];
            if constexpr (FIXED) {
                const float wq = wq_den * ((float)q.w * pw_scale);
                const float az[2] = {(float)du[2] * wq, (float)dl[2] * wq};
                const float byz[4] = {(float)du[1] * az[0], (float)dl[1] * az[0], (float)du[1] * az[1], (float)dl[1] * az[1]};
#pragma unroll
                for (int c = 0; c < 8; ++c) v[c] = T(((c & 1) ? (float)dl[0] : (float)du[0]) * byz[c >> 1]);
            } else {
                const T weight = pose.ow * q.w;
#pragma unroll
                for (int c = 0; c < 8; ++c) v[c] = corner_weight<T, 3>(c, dl, du) * weight;
            }
            uint32_t msum = 0;
            auto add = [&](int o, T val) {
                if constexpr (FIXED) {
                    const uint32_t qv = __float_as_uint((float)val);
                    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(tile_s + (uint32_t)o * 4u), "r"(qv) : "memory");
                    msum += qv;
                } else {
                    atomicAdd(tile + o, val);
                }
            };
            const bool inside = e < n_int && (unsigned)lx < (unsigned)(TX - 1) && (unsigned)ly < (unsigned)(TY - 1) &&
                                (unsigned)lz < (unsigned)(TZ - 1) && i0[0] + 1 < grid.g[0] && i0[1] + 1 < grid.g[1] && i0[2] + 1 < grid.g[2];
            if (inside) {
                const int off = lz * FT::PLANE + ly * FT::PITCH + lx;
#pragma unroll
                for (int c = 0; c < 8; ++c) add(off + (c & 1) + ((c >> 1) & 1) * FT::PITCH + ((c >> 2) & 1) * FT::PLANE, v[c]);
            } else {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int cx = lx + (c & 1), cy = ly + ((c >> 1) & 1), cz = lz + ((c >> 2) & 1);
                    // inside this item's tile and inside the grid (per-corner bounds rule, src/raster.jl:62)
                    const bool in = (unsigned)cx < (unsigned)TX && (unsigned)cy < (unsigned)TY && (unsigned)cz < (unsigned)TZ &&
                                    ox + cx < grid.g[0] && oy + cy < grid.g[1] && oz + cz < grid.g[2];
                    if (in) add(cz * FT::PLANE + cy * FT::PITCH + cx, v[c]);
                }
            }
            mass += msum;       // 8 values below 2^22 each: no 32-bit overflow inside one entry
            }
        }
    };
    if constexpr (sizeof(T) == 4) {
        if (fixed) {
            accumulate(std::true_type{});
            __syncthreads();
            // a cell receives at most `total` contributions of at most 2^F: light tiles cannot wrap
            const bool can_wrap = ((unsigned long long)total << fixed_bits) >= (1ull << 32);
            if (!can_wrap) {
                flush(std::integral_constant<int, 2>{});
                return;
            }
            const long long cells = flush(std::integral_constant<int, 3>{});          // optimistic; zeroes the tile
            if (block_sum_i64(cells - mass, scratch) == 0) return;      // else a wrapped cell: the sums differ, exactly
            __syncthreads();
        }
    }
    accumulate(std::false_type{});
    __syncthreads();
    flush(std::integral_constant<int, 1>{});
}

// ---------------------------------------------------------------------------------------------------------
// 5b. pullback (src/raster_pullback.jl:2-82 per pose; ext/DiffPointRasterisationCUDAExt.jl:19-210 is what it replaces)
//
// USE_TMA: thread 0 fetches the item's ds_dout tile with ONE 4-d tensor-map TMA copy (cp.async.bulk.tensor, SASS UTMALDG;
// cells outside the volume arrive as zeros) as the very first thing the CTA does; the range look-up and the first entry /
// point gathers overlap with it.
//   (Round 1 had declared tensor-map TMA unusable on this pool - "illegal instruction".  The fault is a constraint, not a
//    defect: the box must START on a 16-byte boundary in the innermost dimension (c0 * sizeof(T) % 16 == 0), and both
//    round-1 probes used c0 = -10.  tools/probe_tma_tensor3.cu, profiles/tma_probe_r02.log.  Tile origins are multiples
//    of TX = 32 cells here.)
// !USE_TMA (rows that are not 16-byte multiples): the CTA loads the tile cooperatively.
// d_background (src/raster_pullback.jl:78) is summed from the staged tile (zero outside the volume): ds_dout is read once.
// The sample and its three derivatives come from the factorised trilinear form (25 flops instead of ~150 for the
// corner-by-corner sums).
// ---------------------------------------------------------------------------------------------------------
// The ds_dout tiles are read exactly once: the copy carries an L2 evict-first policy so that the 1.07 GB stream does not push
// the packed gradient buffer (16 MB of red.global targets), the packed points and the entry lists out of L2 (ncu, config 3:
// 1.49 GB of DRAM traffic for 1.09 GB of algorithmic bytes without the hint).
__device__ __forceinline__ void tma_load_tile4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4, %5}], [%6], %7;" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
// Sum 16 values across the warp.  Returns, in lane L, the total of value index L >> 1.
template <typename T>
__device__ __forceinline__ T butterfly16(T (&a)[16], int lane) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {          // lanes with bit 4 keep the upper half of the indices
        const bool hi = lane & 16;
        const T send = hi ? a[i] : a[i + 8], keep = hi ? a[i + 8] : a[i];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const bool hi = lane & 8;
        const T send = hi ? a[i] : a[i + 4], keep = hi ? a[i + 4] : a[i];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const bool hi = lane & 4;
        const T send = hi ? a[i] : a[i + 2], keep = hi ? a[i + 2] : a[i];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    {
        const bool hi = lane & 2;
        const T send = hi ? a[0] : a[1], keep = hi ? a[1] : a[0];
        a[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    return a[0] + __shfl_xor_sync(0xffffffffu, a[0], 1);
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
__global__ void __launch_bounds__(kThreads, resident_ctas(sizeof(T) * kTileCells))
pullback_tile3d_kernel(const __grid_constant__ CUtensorMap map, const T* __restrict__ ds_dout, const Pt4<T>* __restrict__ pts4,
                       const uint32_t* __restrict__ entries, const uint32_t* __restrict__ cnt, const T* __restrict__ rotation,
                       const T* __restrict__ translation, const T* __restrict__ out_weight, Pt4<T>* __restrict__ acc4,
                       T* __restrict__ d_rotation, T* __restrict__ d_translation, T* __restrict__ d_background,
                       T* __restrict__ d_out_weight, Grid<T, 3> grid, TileGeom tg, int64_t b0,
                       unsigned long long* __restrict__ cache_valid) {
    constexpr int NR = 3 * N_IN, NV = NR + 3 + 1;     // d_rotation (col-major 3 x N_IN), d_translation, d_out_weight
    using CM = CellMap<T>;
    if (threadIdx.x == 0 && cache_valid && (blockIdx.x | blockIdx.y | blockIdx.z) == 0) *cache_valid = 1ull;
    constexpr int VEC = CM::VEC;
    struct alignas(16) Pack { T v[VEC]; };
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* tile = reinterpret_cast<T*>(smem_raw);
    __shared__ SlotTable tab;
    __shared__ __align__(8) uint64_t bar;
    __shared__ T red[kThreads / 32][NV + 1];
    // launch grid (nt0, nt1, nt2 * poses): one division instead of four
    const int tx = blockIdx.x, ty = blockIdx.y;
    const int bl = blockIdx.z / tg.nt[2], tz = blockIdx.z - bl * tg.nt[2];
    const int64_t b = b0 + bl;
    const int ox = tx * TX, oy = ty * TY, oz = tz * TZ;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const CM cm;

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
#pragma unroll
        for (int r = 0; r < CM::kPieces; ++r) {
            const int gx = ox + cm.x4, gy = oy + cm.y + CM::dy(r), gz = oz + cm.z0 + CM::dz(r);
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
            reinterpret_cast<Pack*>(tile)[r * kThreads + threadIdx.x] = pk;
        }
    }
    build_slot_table(tab, cnt, tg, bl, tx, ty, tz);
    __syncthreads();                       // table (and the mbarrier init / the cooperative tile load) visible
    const uint32_t total = tab.total;
    if (total == 0 && !d_background) {
        if (USE_TMA) mbar_wait(&bar, 0);   // never leave a copy in flight into shared memory that the next CTA will own
        return;
    }

    // the first entries / points of this thread take off before the tile is waited for.  (Rolling prefetch, two points and
    // three entries ahead: loading kBatch entries and points at once, as the forward does, leaves the memory pipe idle while
    // the batch is computed and cost 27 us on config 3 - the TMA wait already hides the first round trips here.)
    EntryCursor cur;
    uint32_t e = threadIdx.x;
    uint32_t idx = 0, idx_n = 0;
    Pt4<T> qn;
    qn.x = qn.y = qn.z = qn.w = T(0);
    uint32_t idx_nn = 0;
    Pt4<T> qnn = qn;
    if (e < total) {
        cur.init(&tab, entries);
        idx = cur.fetch(e);
        if (e + kThreads < total) idx_n = cur.fetch(e + kThreads);
        if (e + 2 * kThreads < total) idx_nn = cur.fetch(e + 2 * kThreads);
        qn = pts4[idx];
        if (e + kThreads < total) qnn = pts4[idx_n];
    }
    if (USE_TMA) mbar_wait(&bar, 0);

    T bg_part = T(0);
    if (d_background) {                    // src/raster_pullback.jl:78 - from the staged tile (zero outside the volume)
#pragma unroll
        for (int r = 0; r < CM::kPieces; ++r) {
            const Pack pk = reinterpret_cast<const Pack*>(tile)[r * kThreads + threadIdx.x];
#pragma unroll
            for (int k = 0; k < VEC; ++k) bg_part += pk.v[k];
        }
    }
    if (total == 0) {                      // only the tile sum
        bg_part = warp_sum(bg_part);
        if (lane == 0) red[warp][0] = bg_part;
        __syncthreads();
        if (threadIdx.x == 0) {
            T r = T(0);
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w) r += red[w][0];
            red_add(d_background + b, r);
        }
        return;
    }
    T acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = T(0);
    Pose<T, N_IN, 3> pose;
    load_pose<T, N_IN, 3>(pose, rotation, translation, out_weight, b);
    const uint32_t n_int = tab.n_interior;
    for (; e < total; e += kThreads) {
        const Pt4<T> q = qn;
        const uint32_t idx_c = idx;
        idx = idx_n;
        qn = qnn;                                       // two points in flight per thread
        idx_n = idx_nn;
        if (e + 2 * kThreads < total) qnn = pts4[idx_n];
        if (e + 3 * kThreads < total) idx_nn = cur.fetch(e + 3 * kThreads);
        T x[N_IN];
        load_xyz<T, N_IN>(x, q);
        int i0[3];
        T dl[3];
        if (!stencil<T, N_IN, 3>(x, pose, grid, i0, dl)) continue;
        const int lx = i0[0] - ox, ly = i0[1] - oy, lz = i0[2] - oz;
        // corner values in bit order (x = bit 0).  Corners of other tiles are that tile's item's job (everything is
        // linear in G); corners outside the volume read the zero fill = skipped (src/raster_pullback.jl:51)
        T G[8];
        const bool inside = e < n_int && (unsigned)lx < (unsigned)(TX - 1) && (unsigned)ly < (unsigned)(TY - 1) && (unsigned)lz < (unsigned)(TZ - 1);
        if (inside) {
            const int off = (lz * TY + ly) * TX + lx;
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
        for (int jj = 0; jj < N_IN; ++jj) {
            T d = T(0);
#pragma unroll
            for (int n = 0; n < 3; ++n) {
                acc[n + jj * 3] += scaled[n] * x[jj];     // d_rotation, :69
                d += pose.R[n][jj] * scaled[n];           // R' * scaled, :70
            }
            dpt[jj] = d;
        }
        // pose-sum of d_points (:71, :141) and d_point_weight (:58, :146): one 16-byte reduction into the packed,
        // L2-resident buffer (sorted point order)
        red_add4(acc4 + idx_c, dpt[0], dpt[1], dpt[2], s * pose.ow);
    }
    // per-pose sums of this CTA: a transposing butterfly (every shuffle step halves the number of live values: 16 shuffles
    // for 16 values instead of 5 each - the plain reduction was 22 % of this kernel's instructions), one line of shared
    // memory per warp, then one REDG per value
    {
        T vals[16];
#pragma unroll
        for (int v = 0; v < 16; ++v) vals[v] = v < NV ? acc[v] : (v == NV ? bg_part : T(0));
        const T r = butterfly16(vals, lane);            // lane L holds the warp total of value L >> 1 (both lanes of a pair)
        if ((lane & 1) == 0 && (lane >> 1) <= NV) red[warp][lane >> 1] = r;
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

// 6. packed, sorted-order gradients -> d_points (N_in, P) and d_point_weight (P) in the caller's point order: a GATHER through
// the inverse permutation (coalesced stores, 16-byte reads of the L2-resident packed buffer; the scatter through the forward
// permutation wrote 12-byte pieces at random and took 25 us for 1 M points)
template <typename T, int N_IN>
__global__ void __launch_bounds__(256) unpermute_kernel(const Pt4<T>* __restrict__ acc4, const int32_t* __restrict__ inv_perm, int64_t P,
                                                        T* __restrict__ d_points, T* __restrict__ d_point_weight) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += stride) {
        const Pt4<T> a = acc4[__ldg(inv_perm + p)];
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
    size_t off_keys = 0, off_perm = 0, off_pts4 = 0, off_acc4 = 0, off_entries = 0, off_tile_keys = 0, total = 0;
};

// (the layout is the same for the forward and the pullback - the accumulator of the pullback is always reserved - so that
// both can share one workspace and the binning cache)
inline Plan make_plan(int n_in, const int64_t* grid, int64_t P, int64_t B, int sizeof_T) {
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
    if (group > 65535) group = 65535;                                   // gridDim.y of the binning kernels
    if (pl.tg.nt[1] > 65535 || pl.tg.nt[2] > 65535) return pl;          // gridDim.y / gridDim.z of the tile kernels
    if (group * pl.tg.nt[2] > 65535) group = 65535 / pl.tg.nt[2];
    if (group < 1) group = 1;
    if (group * n_tiles > (int64_t)0x7fffffff) return pl;
    pl.group = group;
    // pre-sort cells: ~4 points per cell, 2..6 bits per dimension (measured on B200, 1 M points: 2^15 cells make the
    // histogram and scatter atomics three times slower than 2^18 cells - same-address contention at L2)
    int bits = 2;
    while (bits < 6 && ((int64_t)2 << (n_in * (bits + 1))) <= P) ++bits;
    pl.sort_bits = bits;
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    size_t o = 256;
    pl.off_keys = o;    o = al(o + sizeof(uint32_t) * (size_t)P);
    pl.off_perm = o;    o = al(o + sizeof(int32_t) * (size_t)P);
    pl.off_pts4 = o;    o = al(o + (size_t)sizeof_T * 4 * (size_t)P);
    pl.off_acc4 = o;    o = al(o + (size_t)sizeof_T * 4 * (size_t)P);
    pl.sort_scan = make_scan_region(o, (int64_t)1 << (bits * n_in));
    o = pl.sort_scan.off_ticket + pl.sort_scan.bytes;
    pl.tile_scan = make_scan_region(o, group * n_tiles * 8);
    o = pl.tile_scan.off_ticket + pl.tile_scan.bytes;
    pl.off_entries = o; o = al(o + sizeof(uint32_t) * (size_t)(P * group));
    pl.off_tile_keys = o; o = al(o + sizeof(uint32_t) * (size_t)(P * group));       // keys of pass 1, re-read by pass 2
    pl.total = o;
    pl.ok = true;
    return pl;
}

// per-call state of the binning cache
struct CacheCtl {
    bool enabled = false;
    const unsigned int* skip = nullptr;          // device flag: 1 = the cached bins are valid, binning kernels return
    unsigned long long* valid = nullptr;         // device word the tile kernel sets once the bins are complete
};
inline size_t cache_valid_offset() { return offsetof(CacheHeader, valid); }

// Hashes the inputs and decides (on the device) whether the bins in the workspace can be reused.
template <typename T>
static int cache_begin(CacheCtl& ctl, char* ws, const Plan& pl, int n_in, const int64_t* grid, const T* points, const T* point_weight,
                       const T* rotation, const T* translation, int64_t P, int64_t B, const DeviceInfo& dev, cudaStream_t stream) {
    ctl = CacheCtl{};
    if (tuning().binning_cache != 1) return DPR_OK;
    CacheHeader* h = reinterpret_cast<CacheHeader*>(ws);
    if (pl.group != B) {            // several passes share the entry buffer: nothing to keep; un-mark whatever is there
        DPR_CUDA_TRY(cudaMemsetAsync(ws + cache_valid_offset(), 0, sizeof(unsigned long long), stream));
        return DPR_OK;
    }
    DPR_CUDA_TRY(cudaMemsetAsync(&h->acc[0], 0, sizeof(CacheHeader) - offsetof(CacheHeader, acc), stream));
    unsigned long long params = 0x9e3779b97f4a7c15ull;
    auto fold = [&](unsigned long long v) { params = (params ^ v) * 0xff51afd7ed558ccdull; params ^= params >> 29; };
    fold((unsigned long long)P); fold((unsigned long long)B); fold((unsigned long long)n_in); fold(sizeof(T));
    fold((unsigned long long)grid[0]); fold((unsigned long long)grid[1]); fold((unsigned long long)grid[2]);
    fold(point_weight ? 1ull : 0ull); fold((unsigned long long)pl.total);
    const int64_t n = P * n_in + (point_weight ? P : 0) + B * 3 * n_in + B * 3;
    // few, long-running CTAs: every CTA ends with three atomics on the same line (3000 CTAs made that 15 us)
    int64_t blocks = (n + 256 * 8 - 1) / (256 * 8);
    if (blocks > (int64_t)dev.sm_count * 2) blocks = (int64_t)dev.sm_count * 2;
    if (blocks < 1) blocks = 1;
    {
        LaunchScope scope("tile3_cache_hash", stream);
        hash_inputs_kernel<T><<<(unsigned)blocks, 256, 0, stream>>>(points, P * n_in, point_weight, point_weight ? P : 0, rotation, B * 3 * n_in,
                                                                    translation, B * 3, params, h);
    }
    DPR_CUDA_TRY(cudaGetLastError());
    ctl.enabled = true;
    ctl.skip = &h->skip;
    ctl.valid = &h->valid;
    return DPR_OK;
}

template <typename T, int N_IN>
static int presort(const T* points, const T* point_weight, int64_t P, char* ws, const Plan& pl, const CacheCtl& ctl,
                   const DeviceInfo& dev, cudaStream_t stream) {
    uint32_t* keys = reinterpret_cast<uint32_t*>(ws + pl.off_keys);
    uint32_t* counts = reinterpret_cast<uint32_t*>(ws + pl.sort_scan.off_data);
    int rc;
    if (ctl.enabled) {
        // both scan regions (adjacent) in one conditional pass: a cache hit must keep the counters and the statistics
        const size_t lo = pl.sort_scan.off_ticket, hi = pl.tile_scan.off_ticket + pl.tile_scan.bytes;
        LaunchScope scope("tile3_clear", stream);
        clear_unless_cached_kernel<<<(unsigned)dev.sm_count * 4, 256, 0, stream>>>(reinterpret_cast<uint4*>(ws + lo), (int64_t)((hi - lo) / 16), ctl.skip);
    } else {
        rc = clear_scan_region(ws, pl.sort_scan, stream);
        if (rc != DPR_OK) return rc;
    }
    int64_t blocks = (P + 255) / 256;
    if (blocks > (int64_t)dev.sm_count * 16) blocks = (int64_t)dev.sm_count * 16;
    {
        LaunchScope scope("tile3_sort_count", stream);
        sort_count_kernel<T, N_IN><<<(unsigned)blocks, 256, 0, stream>>>(points, P, pl.sort_bits, keys, counts, ctl.skip);
    }
    rc = launch_scan(ws, pl.sort_scan, stream, "tile3_sort_scan", ctl.skip);
    if (rc != DPR_OK) return rc;
    {
        LaunchScope scope("tile3_sort_scatter", stream);
        sort_scatter4_kernel<T, N_IN><<<(unsigned)blocks, 256, 0, stream>>>(
            points, point_weight, P, keys, counts, reinterpret_cast<int32_t*>(ws + pl.off_perm),
            reinterpret_cast<Pt4<T>*>(ws + pl.off_pts4), reinterpret_cast<uint32_t*>(ws + pl.sort_scan.off_ticket), ctl.skip);
    }
    DPR_CUDA_TRY(cudaGetLastError());
    return DPR_OK;
}

// count + scan + scatter for poses [b0, b0 + nb)
template <typename T, int N_IN>
static int bin_poses(const T* rotation, const T* translation, const Grid<T, 3>& grid, int64_t P, int64_t b0, int64_t nb,
                     char* ws, const Plan& pl, const CacheCtl& ctl, cudaStream_t stream) {
    ScanRegion sr = pl.tile_scan;
    sr.n = nb * pl.tg.n_tiles * 8;
    sr.chunks = (int)((sr.n + kScanChunk - 1) / kScanChunk);
    int rc;
    if (!ctl.enabled) {             // (with the cache the region was cleared, conditionally, together with the pre-sort's)
        rc = clear_scan_region(ws, sr, stream);
        if (rc != DPR_OK) return rc;
    }
    const Pt4<T>* pts4 = reinterpret_cast<const Pt4<T>*>(ws + pl.off_pts4);
    uint32_t* cnt = reinterpret_cast<uint32_t*>(ws + sr.off_data);
    uint32_t* entries = reinterpret_cast<uint32_t*>(ws + pl.off_entries);
    uint32_t* tile_keys = reinterpret_cast<uint32_t*>(ws + pl.off_tile_keys);
    const dim3 gridDim3((unsigned)((P + 1023) / 1024), (unsigned)nb);
    {
        LaunchScope scope("tile3_bin_count", stream);
        tile_count_kernel<T, N_IN><<<gridDim3, 256, 0, stream>>>(pts4, (int)P, rotation, translation, grid, pl.tg, cnt, tile_keys, b0, ctl.skip);
    }
    rc = launch_scan(ws, sr, stream, "tile3_bin_scan", ctl.skip);
    if (rc != DPR_OK) return rc;
    {
        LaunchScope scope("tile3_bin_scatter", stream);
        const dim3 grid_scatter((unsigned)((P + 256 * kScatterK - 1) / (256 * kScatterK)), (unsigned)nb);
        tile_scatter_kernel<<<grid_scatter, 256, 0, stream>>>((int)P, cnt, tile_keys, entries, ctl.skip);
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
