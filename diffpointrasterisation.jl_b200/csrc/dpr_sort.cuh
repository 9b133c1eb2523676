// dpr_sort.cuh - spatial binning of the point cloud (counting sort by the Hilbert index - Z-order for 1 bit per
// dimension - of the cell that holds the point).
//
// The pullback gathers ds_dout at the projected position of every point.  With points in their given (arbitrary)
// order the 32 lanes of a warp hit 32 unrelated 128-byte lines per gather instruction and the L1 data pipe saturates
// (profiles/: 88 % busy, issue slots 52 %).  A rigid pose maps points that are close in space to pixels that are
// close in the image, for EVERY pose, so sorting the points once per call along a space-filling curve makes the lanes of a
// warp land in a small blob of each pose image and share cache lines.  The sort is a three-kernel counting sort
// (histogram, single-CTA scan, scatter); the order inside a bin is arbitrary.  Gradients are written back through
// the permutation, so callers never see the reordering.
#pragma once
#include "dpr_common.cuh"
#include "dpr_internal.h"

namespace dpr {

#ifndef DPR_HILBERT
#define DPR_HILBERT 1
#endif
constexpr bool kUseHilbert = DPR_HILBERT != 0;

__device__ __forceinline__ uint32_t part1by1(uint32_t x) {   // spread the low 16 bits to even positions
    x &= 0x0000ffffu;
    x = (x | (x << 8)) & 0x00ff00ffu;
    x = (x | (x << 4)) & 0x0f0f0f0fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
}
__device__ __forceinline__ uint32_t part1by2(uint32_t x) {   // spread the low 10 bits to every third position
    x &= 0x000003ffu;
    x = (x | (x << 16)) & 0x030000ffu;
    x = (x | (x << 8)) & 0x0300f00fu;
    x = (x | (x << 4)) & 0x030c30c3u;
    x = (x | (x << 2)) & 0x09249249u;
    return x;
}

// Hilbert index of an N-d cell (Skilling's transpose algorithm, "Programming the Hilbert curve", 2004): unlike the
// Z-order curve it has no jumps, so a run of consecutive keys is a compact blob.  Returns the bits*N-bit index.
template <int N>
__device__ __forceinline__ uint32_t hilbert_index(uint32_t (&X)[N], int bits) {
    const uint32_t M = 1u << (bits - 1);
    for (uint32_t Q = M; Q > 1; Q >>= 1) {                 // inverse undo
        const uint32_t P = Q - 1;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (X[i] & Q) X[0] ^= P;
            else { const uint32_t t = (X[0] ^ X[i]) & P; X[0] ^= t; X[i] ^= t; }
        }
    }
#pragma unroll
    for (int i = 1; i < N; ++i) X[i] ^= X[i - 1];         // Gray encode
    uint32_t t = 0;
    for (uint32_t Q = M; Q > 1; Q >>= 1)
        if (X[N - 1] & Q) t ^= Q - 1;
#pragma unroll
    for (int i = 0; i < N; ++i) X[i] ^= t;
    // interleave the transposed form: bit b of X[i] is bit (b * N + (N - 1 - i)) of the index
    uint32_t h = 0;
    for (int b = bits - 1; b >= 0; --b)
#pragma unroll
        for (int i = 0; i < N; ++i) h = (h << 1) | ((X[i] >> b) & 1u);
    return h;
}

template <typename T, int N_IN>
__device__ __forceinline__ uint32_t morton_key(const T* __restrict__ points, int64_t p, int bits) {
    uint32_t q[N_IN];
    const float cells = (float)(1 << bits);
#pragma unroll
    for (int j = 0; j < N_IN; ++j) {
        // nominal cube is (-1, 1); keep a margin and clamp (NaN -> 0)
        float c = ((float)__ldg(points + p * N_IN + j) + 1.25f) * (cells / 2.5f);
        c = fminf(fmaxf(c, 0.f), cells - 1.f);
        q[j] = (uint32_t)c;
    }
    if constexpr (N_IN == 1) return q[0];
    else if (bits >= 2 && kUseHilbert) return hilbert_index<N_IN>(q, bits);
    else if constexpr (N_IN == 2) return part1by1(q[0]) | (part1by1(q[1]) << 1);
    else return part1by2(q[0]) | (part1by2(q[1]) << 1) | (part1by2(q[2]) << 2);
}

template <typename T, int N_IN>
__global__ void __launch_bounds__(256) bin_count_kernel(const T* __restrict__ points, int64_t P, int bits,
                                                        uint32_t* __restrict__ keys, uint32_t* __restrict__ counts) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += stride) {
        const uint32_t k = morton_key<T, N_IN>(points, p, bits);
        keys[p] = k;
        atomicAdd(counts + k, 1u);
    }
}

// exclusive scan of n_bins counters in place (single CTA of 1024 threads).  Each WARP owns a contiguous segment and
// walks it with coalesced 16-byte loads (a per-thread contiguous run made every load touch 32 lines: 117 us for 2^18
// bins); the 32 segment totals are scanned once in between.
static __global__ void __launch_bounds__(1024) bin_scan_kernel(uint32_t* __restrict__ counts, int n_bins) {
    __shared__ uint32_t warp_tot[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (n_bins < 4096 || (n_bins & 4095)) {
        // small or odd sizes: one thread per contiguous run (a few hundred elements in total)
        const int per = (n_bins + 1023) / 1024;
        const int lo = threadIdx.x * per, hi = (lo + per < n_bins) ? lo + per : n_bins;
        uint32_t sum = 0;
        for (int i = lo; i < hi; ++i) sum += counts[i];
        uint32_t incl = sum;
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
        uint32_t run = (warp ? warp_tot[warp - 1] : 0u) + incl - sum;
        for (int i = lo; i < hi; ++i) {
            const uint32_t v = counts[i];
            counts[i] = run;
            run += v;
        }
        return;
    }
    const int seg = n_bins / 32;                      // elements per warp, a multiple of 128
    uint4* v4 = reinterpret_cast<uint4*>(counts + (size_t)warp * seg);
    const int rounds = seg / 128;
    uint32_t tot = 0;
#pragma unroll 4
    for (int r = 0; r < rounds; ++r) {
        const uint4 q = v4[r * 32 + lane];
        tot += q.x + q.y + q.z + q.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (lane == 0) warp_tot[warp] = tot;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warp_tot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        warp_tot[lane] = w;                           // inclusive over warps
    }
    __syncthreads();
    uint32_t run = warp ? warp_tot[warp - 1] : 0u;
    for (int r = 0; r < rounds; ++r) {
        const uint4 q = v4[r * 32 + lane];
        const uint32_t s = q.x + q.y + q.z + q.w;
        uint32_t incl = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        uint4 o4;
        o4.x = run + incl - s; o4.y = o4.x + q.x; o4.z = o4.y + q.y; o4.w = o4.z + q.z;
        v4[r * 32 + lane] = o4;
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
}

// `spread`: scatter forward-pass style - inside every full run of 1024 sorted points the order is permuted with a
// stride permutation (i -> 33 i mod 1024), so the 32 lanes of a warp hold points that are ~32 apart in the sorted
// order: the run stays spatially compact (culling, bounding boxes) but neighbouring lanes no longer hit the same cell
// with their shared-memory atomics.
template <typename T, int N_IN>
__global__ void __launch_bounds__(256) bin_scatter_kernel(const T* __restrict__ points, const T* __restrict__ point_weight,
                                                          int64_t P, const uint32_t* __restrict__ keys,
                                                          uint32_t* __restrict__ offsets, int32_t* __restrict__ perm,
                                                          T* __restrict__ sorted_points, T* __restrict__ sorted_pw,
                                                          int spread) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const uint32_t full_runs_end = (uint32_t)(P & ~(int64_t)1023);
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += stride) {
        uint32_t pos = atomicAdd(offsets + keys[p], 1u);
        if (spread && pos < full_runs_end) pos = (pos & ~1023u) | (((pos & 1023u) * 33u) & 1023u);
        perm[pos] = (int32_t)p;
#pragma unroll
        for (int j = 0; j < N_IN; ++j) sorted_points[(int64_t)pos * N_IN + j] = __ldg(points + p * N_IN + j);
        if (point_weight) sorted_pw[pos] = __ldg(point_weight + p);
    }
}

// Axis-aligned bounding box (centre, half extent) of every run of `chunk` consecutive (sorted) points: lets a CTA
// skip whole runs that cannot touch its tile.  One CTA of 256 threads per run.
template <typename T, int N_IN>
__global__ void __launch_bounds__(256) chunk_aabb_kernel(const T* __restrict__ points, int64_t P, int chunk,
                                                         float* __restrict__ aabb) {
    const int64_t lo = (int64_t)blockIdx.x * chunk;
    const int64_t hi = (lo + chunk < P) ? lo + chunk : P;
    float mn[N_IN], mx[N_IN];
#pragma unroll
    for (int j = 0; j < N_IN; ++j) { mn[j] = 3.4e38f; mx[j] = -3.4e38f; }
    for (int64_t p = lo + threadIdx.x; p < hi; p += blockDim.x) {
#pragma unroll
        for (int j = 0; j < N_IN; ++j) {
            const float v = (float)__ldg(points + p * N_IN + j);
            mn[j] = fminf(mn[j], v);
            mx[j] = fmaxf(mx[j], v);
        }
    }
    __shared__ float smn[8][N_IN], smx[8][N_IN];
#pragma unroll
    for (int j = 0; j < N_IN; ++j) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[j] = fminf(mn[j], __shfl_xor_sync(0xffffffffu, mn[j], o));
            mx[j] = fmaxf(mx[j], __shfl_xor_sync(0xffffffffu, mx[j], o));
        }
        if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5][j] = mn[j]; smx[threadIdx.x >> 5][j] = mx[j]; }
    }
    __syncthreads();
    if (threadIdx.x < N_IN) {
        const int j = threadIdx.x;
        float a = smn[0][j], b = smx[0][j];
        for (int w = 1; w < 8; ++w) { a = fminf(a, smn[w][j]); b = fmaxf(b, smx[w][j]); }
        // widen a little: the Float64 points were rounded to float for the box
        const float ctr = 0.5f * (a + b), half = 0.5f * (b - a) * 1.0001f + 1e-6f * (fabsf(a) + fabsf(b));
        aabb[(int64_t)blockIdx.x * 2 * N_IN + j] = ctr;
        aabb[(int64_t)blockIdx.x * 2 * N_IN + N_IN + j] = half;
    }
}

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
                                                                    uint32_t* __restrict__ ticket, const unsigned int* __restrict__ skip) {
    __shared__ uint32_t s_chunk, s_base, warp_tot[32];
    if (skip && *skip) return;                 // binning cache hit (uniform)
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
static int launch_scan(char* ws, const ScanRegion& r, cudaStream_t stream, const char* name, const unsigned int* skip = nullptr) {
    LaunchScope scope(name, stream);
    scan_lookback_kernel<<<(unsigned)r.chunks, 1024, 0, stream>>>(reinterpret_cast<uint32_t*>(ws + r.off_data), r.n,
                                                                  reinterpret_cast<unsigned long long*>(ws + r.off_state),
                                                                  reinterpret_cast<uint32_t*>(ws + r.off_ticket), skip);
    DPR_CUDA_TRY(cudaGetLastError());
    return DPR_OK;
}

struct SortPlan {
    int bits = 0;             // bits per dimension
    int n_bins = 0;
    size_t off_keys = 0, off_perm = 0, off_points = 0, off_pw = 0, total = 0;
    ScanRegion scan;          // histogram / bin offsets (scan.off_data) with the scan's ticket and state words in front
};

inline SortPlan make_sort_plan(int n_in, int64_t P, int sizeof_T, bool has_pw, size_t base_offset) {
    SortPlan sp;
    // ~2 points per bin on average: finer bins do not make a warp's 32 points any more compact
    int total_bits = 6;
    while (total_bits < 18 && ((int64_t)2 << total_bits) < P) ++total_bits;
    sp.bits = total_bits / n_in;
    if (sp.bits < 1) sp.bits = 1;
    if (sp.bits > 9) sp.bits = 9;
    sp.n_bins = 1 << (sp.bits * n_in);
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    size_t o = al(base_offset);
    sp.off_keys = o;   o = al(o + sizeof(uint32_t) * (size_t)P);
    sp.scan = make_scan_region(o, sp.n_bins);
    o = al(o + sp.scan.bytes);
    sp.off_perm = o;   o = al(o + sizeof(int32_t) * (size_t)P);
    sp.off_points = o; o = al(o + (size_t)sizeof_T * (size_t)P * n_in);
    sp.off_pw = o;     o = al(o + (has_pw ? (size_t)sizeof_T * (size_t)P : 0));
    sp.total = o;
    return sp;
}

template <typename T, int N_IN>
static int sort_points(const T* points, const T* point_weight, int64_t P, void* workspace, const SortPlan& sp,
                       const DeviceInfo& dev, cudaStream_t stream, bool spread = false) {
    char* ws = static_cast<char*>(workspace);
    uint32_t* keys = reinterpret_cast<uint32_t*>(ws + sp.off_keys);
    uint32_t* counts = reinterpret_cast<uint32_t*>(ws + sp.scan.off_data);
    int32_t* perm = reinterpret_cast<int32_t*>(ws + sp.off_perm);
    T* spts = reinterpret_cast<T*>(ws + sp.off_points);
    T* spw = reinterpret_cast<T*>(ws + sp.off_pw);
    int rc = clear_scan_region(ws, sp.scan, stream);
    if (rc != DPR_OK) return rc;
    int64_t blocks = (P + 255) / 256;
    if (blocks > (int64_t)dev.sm_count * 16) blocks = (int64_t)dev.sm_count * 16;
    {
        LaunchScope scope("bin_count", stream);
        bin_count_kernel<T, N_IN><<<(unsigned)blocks, 256, 0, stream>>>(points, P, sp.bits, keys, counts);
    }
    rc = launch_scan(ws, sp.scan, stream, "bin_scan");      // single pass, one CTA per 4096 bins (one CTA took 35 us for 2^18)
    if (rc != DPR_OK) return rc;
    {
        LaunchScope scope("bin_scatter", stream);
        bin_scatter_kernel<T, N_IN><<<(unsigned)blocks, 256, 0, stream>>>(points, point_weight, P, keys, counts, perm, spts, spw,
                                                                          spread ? 1 : 0);
    }
    DPR_CUDA_TRY(cudaGetLastError());
    return DPR_OK;
}

}  // namespace dpr
