// dpr_forward_radial.cuh - Float32 2-d forward splat, one shared-memory slab per pose, points sorted by RADIUS
// (included by dpr_forward.cu after dpr_forward_fast.cuh).
//
// The ncu captures of fwd_tile2d_fast_kernel on config 2 (profiles/ncu_full_r01_v3_cfg2_summary.csv) show 114 warp
// instructions per 32 splats although the straight-line interior path is ~70: the rest is the machinery for the few
// lanes whose stencil leaves the slab (interior test, ballot, per-warp queue), which almost every warp iteration pays
// because ~5 % of lanes are not interior.  This kernel removes that machinery for most of the points:
//
//   * For a pose with rows r_0, r_1 a point x lands at  coord_k = (r_k . x - origin_k) * scale_k,  and
//     |r_k . x| <= |r_k| |x| (Cauchy-Schwarz).  So every point with |x| <= r_safe(pose) has all four corners inside
//     the slab, where r_safe is the distance from the projected origin to the nearest slab edge divided by
//     |r_k| scale_k.  No assumption that R is a rotation: the row norms are computed from the matrix.
//   * The points are counting-sorted ONCE per call by |x| (radial bins; the order inside a bin is arbitrary) into a
//     packed float4 copy {x, y, z, point_weight} padded to whole 1024-point chunks with NaN points, and the
//     largest radius of every chunk is recorded.  The forward image does not depend on the order of the points.
//   * Per pose, the leading chunks with rmax <= r_safe are processed by a straight-line loop: one LDG.128 per point,
//     no bounds / slab / activity test, no ballot, no queue (phase A, ~80 % of config 2's splats).  The remaining
//     chunks go through the checked loop of the fast kernel (phase B).
//   * Hybrid mode (image a little larger than shared memory): the band of rows kept on chip is centred on the
//     projected origin of EACH pose instead of the image centre, which is where the cloud is.
//
// Accumulation is fixed point on the native ATOMS.ADD with the 64-bit mass checksum and the float CAS fallback of
// dpr_forward_fast.cuh, but the integers are produced by a subnormal product instead of a conversion (see the kernel)
// and the tile is flushed in one pass (checksum + conversion + store from one conflict-free 16-byte read).
// Measured on config 2 (B200): 114 -> 69 instructions per warp-splat, 1.73 -> 1.57 ms; the kernel is then bound by the
// shared-memory data pipe (3.95 wavefronts per ATOMS.ADD from random bank conflicts), see DESIGN.md 4.1b.
#pragma once
#include "dpr_common.cuh"
#include "dpr_sort.cuh"

namespace dpr {

constexpr int kRadialBins = 2048;
constexpr float kRadialMax = 2.0f;       // radii beyond this (and NaN) share the last bin
constexpr int kRadialPad = 3 * kChunk;   // far-away points behind the last chunk: targets of the register prefetch

struct RadialPlan {
    int64_t P_pad = 0;
    int n_chunks = 0;
    size_t off_keys = 0, off_counts = 0, off_rmax = 0, off_pts4 = 0, total = 0;
    size_t zero_bytes = 0;     // counts and rmax are contiguous: one memset
};

inline RadialPlan make_radial_plan(int64_t P, size_t base_offset) {
    RadialPlan rp;
    rp.P_pad = (P + kChunk - 1) / kChunk * kChunk;
    rp.n_chunks = (int)(rp.P_pad / kChunk);
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    size_t o = al(base_offset);
    rp.off_keys = o;   o = al(o + sizeof(uint32_t) * (size_t)P);
    rp.off_counts = o; o = o + sizeof(uint32_t) * (size_t)kRadialBins;
    rp.off_rmax = o;   o = al(o + sizeof(uint32_t) * (size_t)rp.n_chunks);
    rp.zero_bytes = o - rp.off_counts;
    rp.off_pts4 = o;   o = al(o + sizeof(float4) * (size_t)(rp.P_pad + kRadialPad));
    rp.total = o;
    return rp;
}

template <int N_IN>
__device__ __forceinline__ float point_radius(const float* __restrict__ points, int64_t p, float (&x)[N_IN]) {
    float r2 = 0.f;
#pragma unroll
    for (int j = 0; j < N_IN; ++j) { x[j] = __ldg(points + p * N_IN + j); r2 = fmaf(x[j], x[j], r2); }
    return sqrtf(r2);
}
__device__ __forceinline__ uint32_t radial_key(float r) {
    // NaN compares false -> last bin
    return (r < kRadialMax) ? (uint32_t)(r * ((float)(kRadialBins - 1) / kRadialMax)) : (uint32_t)(kRadialBins - 1);
}

template <int N_IN>
__global__ void __launch_bounds__(256) radial_count_kernel(const float* __restrict__ points, int64_t P,
                                                           uint32_t* __restrict__ keys, uint32_t* __restrict__ counts) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += stride) {
        float x[N_IN];
        const uint32_t k = radial_key(point_radius<N_IN>(points, p, x));
        keys[p] = k;
        atomicAdd(counts + k, 1u);
    }
}

// scatter into the packed copy; rmax[chunk] = largest radius of the chunk (as float bits: non-negative floats order
// like unsigned integers, NaN above everything); positions P .. P_pad+kRadialPad-1 are filled with NaN points
// (those behind P_pad are only ever prefetch targets).
template <int N_IN>
__global__ void __launch_bounds__(256) radial_scatter_kernel(const float* __restrict__ points,
                                                             const float* __restrict__ point_weight, int64_t P, int64_t P_pad,
                                                             const uint32_t* __restrict__ keys, uint32_t* __restrict__ offsets,
                                                             float4* __restrict__ pts4, uint32_t* __restrict__ rmax,
                                                             const float* __restrict__ pw_stats) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int em = point_weight_exponent(point_weight ? pw_stats : nullptr);
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P_pad + kRadialPad; p += stride) {
        if (p < P) {
            float x[N_IN];
            const float r = point_radius<N_IN>(points, p, x);
            const uint32_t pos = atomicAdd(offsets + keys[p], 1u);
            float4 v;
            v.x = x[0];
            v.y = N_IN > 1 ? x[N_IN > 1 ? 1 : 0] : 0.f;
            v.z = N_IN > 2 ? x[N_IN > 2 ? 2 : 0] : 0.f;
            v.w = point_weight ? ldexpf(__ldg(point_weight + p), -em) : 1.f;
            pts4[pos] = v;
            atomicMax(rmax + (pos >> 10), __float_as_uint(fabsf(r)));
        } else {
            // NaN coordinates: no in-bounds corner for ANY pose matrix (a far-away point would still land on the
            // projected origin of an all-zero matrix), and never interior because their chunk has rmax = +inf
            const float qnan = __int_as_float(0x7fc00000);
            pts4[p] = make_float4(qnan, qnan, qnan, 0.f);
            if (p < P_pad) atomicMax(rmax + (p >> 10), 0x7f800000u);   // +inf: never a safe chunk
        }
    }
}

template <int N_IN>
static int radial_sort_points(const float* points, const float* point_weight, const float* pw_stats, int64_t P,
                              void* workspace, const RadialPlan& rp, const DeviceInfo& dev, cudaStream_t stream) {
    char* ws = static_cast<char*>(workspace);
    uint32_t* keys = reinterpret_cast<uint32_t*>(ws + rp.off_keys);
    uint32_t* counts = reinterpret_cast<uint32_t*>(ws + rp.off_counts);
    uint32_t* rmax = reinterpret_cast<uint32_t*>(ws + rp.off_rmax);
    float4* pts4 = reinterpret_cast<float4*>(ws + rp.off_pts4);
    DPR_CUDA_TRY(cudaMemsetAsync(counts, 0, rp.zero_bytes, stream));
    int64_t blocks = (rp.P_pad + kRadialPad + 255) / 256;
    if (blocks > (int64_t)dev.sm_count * 16) blocks = (int64_t)dev.sm_count * 16;
    {
        LaunchScope scope("radial_count", stream);
        radial_count_kernel<N_IN><<<(unsigned)blocks, 256, 0, stream>>>(points, P, keys, counts);
    }
    {
        LaunchScope scope("bin_scan", stream);
        bin_scan_kernel<<<1, 1024, 0, stream>>>(counts, kRadialBins);
    }
    {
        LaunchScope scope("radial_scatter", stream);
        radial_scatter_kernel<N_IN><<<(unsigned)blocks, 256, 0, stream>>>(points, point_weight, P, rp.P_pad, keys, counts, pts4, rmax, pw_stats);
    }
    DPR_CUDA_TRY(cudaGetLastError());
    return DPR_OK;
}

// streaming 16-byte load that does not allocate in L1: the packed points are read once per CTA, and an L1 fill costs
// the same data-pipe cycles as the read itself - cycles the shared-memory atomics need (ncu: LSU data pipe 78 % busy)
__device__ __forceinline__ float4 ldg_stream4(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

template <int N_IN, bool HAS_PW>
__global__ void __launch_bounds__(1024, 1)
fwd_tile2d_radial_kernel(const float4* __restrict__ pts4, const float* __restrict__ rmax,
                         const float* __restrict__ rotation, const float* __restrict__ translation,
                         const float* __restrict__ background, const float* __restrict__ out_weight,
                         float* __restrict__ out, Grid<float, 2> grid, int n_chunks_total, FastTileParams tp) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ long long scratch[32];
    __shared__ int s_first;
    const int64_t b = blockIdx.x / tp.splits;
    const int q = blockIdx.x % tp.splits;
    const int g0 = grid.g[0], g1 = grid.g[1];
    const int tile_cap = tp.rows * g0;
    unsigned* tile_u = reinterpret_cast<unsigned*>(smem_raw);
    float* tile_f = reinterpret_cast<float*>(smem_raw);
    int* queue = reinterpret_cast<int*>(smem_raw + (((size_t)tile_cap * 4 + 127) / 128) * 128);
    float* __restrict__ img = out + b * grid.cells;
    const float bg = background ? __ldg(background + b) : 0.f;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int* wq = queue + warp * kQueueCap;

    Pose<float, N_IN, 2> pose;
    load_pose(pose, rotation, translation, out_weight, b);

    // ---- rows kept on chip: the whole image, or (hybrid) tp.rows rows centred on this pose's projected origin ------
    const bool border = tp.rows < g1;
    int band_lo = 0, band_hi = g1;
    if (border) {
        const float cyf = -pose.origin[1] * grid.scale[1];
        const int c = (fabsf(cyf) < 1e9f) ? __float2int_rn(cyf) : g1 / 2;      // NaN -> image centre
        int lo = c - tp.rows / 2;
        lo = lo < 0 ? 0 : lo;
        lo = lo > g1 - tp.rows ? g1 - tp.rows : lo;
        band_lo = lo;
        band_hi = lo + tp.rows;
    }
    const int ys = band_lo, ye = band_hi, nrows = ye - ys, n_tile = nrows * g0;

    const int chunks_per_split = tp.per_split / kChunk;
    const int c_begin = q * chunks_per_split;
    const int c_end = (c_begin + chunks_per_split < n_chunks_total) ? c_begin + chunks_per_split : n_chunks_total;
    const int p_begin = c_begin * kChunk, p_end = c_end * kChunk;

    // The first two points of this thread and its first chunk radius depend on nothing but the block index: they are
    // requested before the pose is loaded, the tile zeroed and the safe radius derived, so those ~2 us of set-up (one CTA
    // per SM: nothing else hides them) overlap with the loads' latency (config 2: 1.572 -> 1.539 ms).  The copy is padded, so
    // the prefetch never needs a bounds test (a split that starts beyond the last chunk - (Q-1)*ceil(n/Q) can exceed n - has
    // no work: its prefetch is clamped into the padding, which covers chunks n .. n+2).
    // (A persistent variant - one CTA per SM walking the poses, the next pose's parameters staged with cp.async - measured
    // 1.548 ms: the hardware's dynamic CTA scheduling balances the poses' unequal outer shells better than a static walk.)
    const float4* __restrict__ next = pts4 + (size_t)(c_begin < n_chunks_total ? c_begin : n_chunks_total) * kChunk + threadIdx.x;
    float4 pf0 = ldg_stream4(next), pf1 = ldg_stream4(next + kChunk);
    next += 2 * kChunk;
    float rmax_first = 0.f;
    if (c_begin + (int)threadIdx.x < c_end) rmax_first = __ldg(rmax + c_begin + threadIdx.x);

    // ---- fixed-point scale (uniform across the CTA); not eligible -> float CAS accumulation (generic code) -----------
    // A contribution v = w * out_weight * point_weight = w * pw' * A with pw' = point_weight * 2^-em in [0, 1] (stored in
    // the packed copy) and A = out_weight * 2^em is accumulated as the integer rint(w * pw' * Q), Q = rint(A * 2^(F - e)),
    // 2^e >= A, so 2^(F-1) < Q <= 2^F.  The integer is produced WITHOUT a conversion: multiplying by Q * 2^-149 (the
    // subnormal whose bit pattern is the integer Q) lands the product in the subnormal range, where round-to-nearest
    // leaves exactly rint(.) in the mantissa bits - __float_as_int of the product IS the fixed-point value.  The
    // rounding of Q is undone exactly at the flush (inv_q = A / Q).
    bool fixed = false;
    float wq_den = 0.f, inv_q = 0.f;
    const int em = point_weight_exponent(HAS_PW ? tp.pw_stats : nullptr);
    {
        bool ok = pose.ow > 0.f;
        if (HAS_PW) {
            const float wmax = __ldg(tp.pw_stats), wmin = __ldg(tp.pw_stats + 1), wmean = __ldg(tp.pw_stats + 2);
            ok = ok && wmin >= 0.f && wmax > 0.f && wmax < 3e38f && wmax <= 64.f * wmean;
        }
        const float A = ldexpf(pose.ow, em);
        ok = ok && A > 1e-30f && A < 1e30f;
        if (ok) {
            int e;
            frexpf(A, &e);                                     // A <= 2^e
            const float Q = rintf(ldexpf(A, tp.fixed_bits - e));   // 2^(F-1) <= Q <= 2^F <= 2^22
            wq_den = __int_as_float((int)Q);
            inv_q = A / Q;
            fixed = true;
        }
    }

    // ---- safe radius of this pose (see the header comment); a NaN anywhere makes every comparison below false ----
    float r_safe;
    {
        float n0 = 0.f, n1 = 0.f;
#pragma unroll
        for (int j = 0; j < N_IN; ++j) { n0 = fmaf(pose.R[0][j], pose.R[0][j], n0); n1 = fmaf(pose.R[1][j], pose.R[1][j], n1); }
        n0 = sqrtf(n0) * grid.scale[0];
        n1 = sqrtf(n1) * grid.scale[1];
        const float cx = -pose.origin[0] * grid.scale[0], cy = -pose.origin[1] * grid.scale[1];
        // interior  <=>  0.5 < coord_x <= g0 - 0.5  and  ys + 0.5 < coord_y <= ye - 0.5; keep a 0.05-cell margin for the
        // rounding of the (unfused, Float32) coordinate arithmetic
        const float mx = fminf(cx - 0.5f, (float)g0 - 0.5f - cx) - 0.05f - 1e-5f * (float)g0;
        const float my = fminf(cy - ((float)ys + 0.5f), ((float)ye - 0.5f) - cy) - 0.05f - 1e-5f * (float)g1;
        r_safe = fminf(mx / fmaxf(n0, 1e-30f), my / fmaxf(n1, 1e-30f)) * 0.9999f;
        r_safe = fminf(r_safe, 1e30f);              // stays below the +inf that marks padding / non-finite chunks
        if (!(mx > 0.f) || !(my > 0.f) || !(n0 < 1e30f) || !(n1 < 1e30f)) r_safe = -1.f;
    }

    if (threadIdx.x == 0) s_first = c_end;
    if ((n_tile & 3) == 0) {
        for (int i = threadIdx.x; i < n_tile / 4; i += blockDim.x) reinterpret_cast<uint4*>(tile_u)[i] = make_uint4(0u, 0u, 0u, 0u);
    } else {
        for (int i = threadIdx.x; i < n_tile; i += blockDim.x) tile_u[i] = 0u;
    }
    if (tp.exclusive && border) {
        const int lo_cells = band_lo * g0;
        for (int i = threadIdx.x; i < lo_cells; i += blockDim.x) img[i] = bg;
        for (int i = band_hi * g0 + threadIdx.x; i < g0 * g1; i += blockDim.x) img[i] = bg;
        __threadfence();
    }
    __syncthreads();
    if (fixed) {
        for (int c = c_begin + (int)threadIdx.x; c < c_end; c += blockDim.x) {
            const float rm = c == c_begin + (int)threadIdx.x ? rmax_first : __ldg(rmax + c);
            if (!(rm <= r_safe)) atomicMin(&s_first, c);
        }
    }
    __syncthreads();
    const int c_safe = s_first;               // chunks [c_begin, c_safe) are interior for this pose

    auto load4 = [&](int p, float (&x)[N_IN], float& pw) {
        const float4 v = __ldg(pts4 + p);
        x[0] = v.x;
        if constexpr (N_IN > 1) x[1] = v.y;
        if constexpr (N_IN > 2) x[2] = v.z;
        if (HAS_PW) pw = ldexpf(v.w, em);      // undo the storage scaling (exact)
    };

    long long mass = 0;
    bool flushed = false;
    if (fixed) {
        float2 Rj[N_IN];
#pragma unroll
        for (int j = 0; j < N_IN; ++j) Rj[j] = make_float2(pose.R[0][j], pose.R[1][j]);
        const float2 neg_origin = make_float2(-pose.origin[0], -pose.origin[1]);
        const float2 scale2 = make_float2(grid.scale[0], grid.scale[1]);
        // 32-bit shared address of cell (0, ys) moved by the "-1"s of the 1-based cell index the stencil produces
        const uint32_t tile_s = smem_u32(tile_u);
        uint32_t base_s = tile_s - (uint32_t)(((ys + 1) * g0 + 1) * 4);
        asm volatile("" : "+r"(base_s));                    // keep it ONE register: do not re-derive it per point
        const uint32_t row_s = (uint32_t)g0 * 4u;
        unsigned mass32 = 0;
        int wq_count = 0;                                   // warp-uniform

        // transform + stencil with both output dimensions packed (see dpr_forward_fast.cuh for the ptxas note: products
        // packed, every add that consumes a product scalar, so no FFMA2 can appear and `coord` is bit-exact).
        // Returns the 1-BASED cell (ref of src/raster.jl:94) and (du, dl) per dimension as adjacent pairs.
        auto stencil_packed = [&](const float4& v, int& rx, int& ry, float2& wx, float2& wy) {
            float xs[3] = {v.x, v.y, v.z};
            float2 prod[N_IN];
#pragma unroll
            for (int j = 0; j < N_IN; ++j) prod[j] = __fmul2_rn(Rj[j], make_float2(xs[j], xs[j]));
            float s0 = prod[0].x, s1 = prod[0].y;
#pragma unroll
            for (int j = 1; j < N_IN; ++j) { s0 = __fadd_rn(s0, prod[j].x); s1 = __fadd_rn(s1, prod[j].y); }
            const float2 coord = __fmul2_rn(__fadd2_rn(make_float2(s0, s1), neg_origin), scale2);
            const float2 r = make_float2(ceilf(__fadd_rn(coord.x, -0.5f)), ceilf(__fadd_rn(coord.y, -0.5f)));
            const float2 t = __fadd2_rn(r, make_float2(-0.5f, -0.5f));
            const float dlx = __fsub_rn(coord.x, t.x), dly = __fsub_rn(coord.y, t.y);   // coord - (r - 0.5)
            wx = make_float2(1.f - dlx, dlx);               // (weight of the lower, of the upper cell) in x
            wy = make_float2(1.f - dly, dly);
            rx = __float2int_rn(r.x);
            ry = __float2int_rn(r.y);
        };
        // the four quantised corner contributions of one interior point: three packed multiplies, no conversion
        auto splat4 = [&](uint32_t addr, const float2& wx, const float2& wy, float wqd) {
            const float2 ab = __fmul2_rn(wy, make_float2(wqd, wqd));
            const float2 q0 = __fmul2_rn(wx, make_float2(ab.x, ab.x));
            const float2 q1 = __fmul2_rn(wx, make_float2(ab.y, ab.y));
            const uint32_t i00 = __float_as_uint(q0.x), i10 = __float_as_uint(q0.y);
            const uint32_t i01 = __float_as_uint(q1.x), i11 = __float_as_uint(q1.y);
            red_shared_u32(addr, i00);
            red_shared_u32(addr + 4u, i10);
            red_shared_u32(addr + row_s, i01);
            red_shared_u32(addr + row_s + 4u, i11);
            mass32 += (i00 + i10) + (i01 + i11);
        };
        // generic (any position) handling of one point: used for the deferred lanes of phase B
        auto slow_point = [&](int p) {
            float x[N_IN], pw = 1.f;
            load4(p, x, pw);
            int i0[2];
            float dl[2];
            if (!stencil(x, pose, grid, i0, dl)) return;
            const float du0 = 1.f - dl[0], du1 = 1.f - dl[1];
            const float w[4] = {du0 * du1, dl[0] * du1, du0 * dl[1], dl[0] * dl[1]};
            const float wqd = HAS_PW ? wq_den * ldexpf(pw, -em) : wq_den;
            const bool x_lo = i0[0] >= 0, x_hi = i0[0] + 1 < g0;
#pragma unroll
            for (int cy = 0; cy < 2; ++cy) {
                const int iy = i0[1] + cy;
                if (iy < 0 || iy >= g1) continue;
                if (iy >= ys && iy < ye) {
                    const uint32_t addr = tile_s + (uint32_t)(((iy - ys) * g0 + i0[0]) * 4);
                    if (x_lo) { const uint32_t v = __float_as_uint(w[2 * cy] * wqd); red_shared_u32(addr, v); mass32 += v; }
                    if (x_hi) { const uint32_t v = __float_as_uint(w[2 * cy + 1] * wqd); red_shared_u32(addr + 4u, v); mass32 += v; }
                } else {
                    // a row outside the on-chip band (hybrid mode only): accumulate in L2
                    const float weight = pose.ow * pw;
                    float* addr = img + (int64_t)iy * g0 + i0[0];
                    const float va = w[2 * cy] * weight, vb = w[2 * cy + 1] * weight;
                    if (x_lo && x_hi) red_add2(addr, va, vb);
                    else if (x_lo) red_add(addr, va);
                    else if (x_hi) red_add(addr + 1, vb);
                }
            }
        };

        // two points in flight per thread (the points stream from L2: one chunk ahead left ~27 % of the stall cycles on
        // the load's scoreboard); pf0, pf1 were requested at the top of the kernel
        auto pop = [&]() -> float4 {
            const float4 v = pf0;
            pf0 = pf1;
            pf1 = ldg_stream4(next);
            next += kChunk;
            return v;
        };
        // ---- phase A: chunks whose points are all interior for this pose --------------------------------------
        for (int it0 = c_begin; it0 < c_safe; it0 += 64) {      // the 32-bit mass partial is folded every 64 chunks
            const int it1 = (it0 + 64 < c_safe) ? it0 + 64 : c_safe;
#pragma unroll 2
            for (int c = it0; c < it1; ++c) {
                const float4 v = pop();
                int rx, ry;
                float2 wx, wy;
                stencil_packed(v, rx, ry, wx, wy);
                splat4(base_s + (uint32_t)((ry * g0 + rx) * 4), wx, wy, HAS_PW ? wq_den * v.w : wq_den);
            }
            mass += mass32;
            mass32 = 0;
        }
        // ---- phase B: the outer shells - interior test, deferred lanes compacted into a per-warp queue ----------
        for (int it0 = c_safe; it0 < c_end; it0 += 64) {
            const int it1 = (it0 + 64 < c_end) ? it0 + 64 : c_end;
            for (int c = it0; c < it1; ++c) {
                const float4 v = pop();
                int rx, ry;
                float2 wx, wy;
                stencil_packed(v, rx, ry, wx, wy);
                // all four corners on chip  <=>  1 <= rx <= g0 - 1  and  ys + 1 <= ry <= ye - 1
                const bool interior = (unsigned)(rx - 1) < (unsigned)(g0 - 1) && (unsigned)(ry - 1 - ys) < (unsigned)(nrows - 1);
                if (interior) splat4(base_s + (uint32_t)((ry * g0 + rx) * 4), wx, wy, HAS_PW ? wq_den * v.w : wq_den);
                // everything else (slab edge, border rows, outside the image, padding) is sorted out by slow_point
                const unsigned slow_mask = __ballot_sync(0xffffffffu, !interior);
                if (slow_mask) {
                    if (!interior) wq[wq_count + __popc(slow_mask & ((1u << lane) - 1u))] = c * kChunk + (int)threadIdx.x;
                    wq_count += __popc(slow_mask);
                    __syncwarp();
                    if (wq_count >= 32) {
                        slow_point(wq[wq_count - 32 + lane]);
                        wq_count -= 32;
                        __syncwarp();
                    }
                }
            }
            mass += mass32;
            mass32 = 0;
        }
        if (lane < wq_count) slow_point(wq[lane]);
        mass += mass32;

        // ---- wrap check and flush -----------------------------------------------------------------------------------
        // The CTA owns its cells (exclusive) in the common case: read the tile ONCE with conflict-free 16-byte loads,
        // summing the cells for the checksum and storing the converted values optimistically; the rare wrapped slab is
        // redone in float below and simply stored again.  (Scalar reads of 4 consecutive cells per thread were 4-way
        // bank conflicts: 51 M shared-load wavefronts per launch in the ncu capture.)
        float* __restrict__ dst0 = img + (int64_t)ys * g0;
        const bool vec_ok = tp.exclusive && (n_tile & 3) == 0 && (reinterpret_cast<uintptr_t>(dst0) & 15) == 0;
        unsigned long long cells_sum = 0;
        __syncthreads();                                      // ends the accumulation
        if (vec_ok) {
            const uint4* t4 = reinterpret_cast<const uint4*>(tile_u);
            for (int i = threadIdx.x; i < n_tile / 4; i += blockDim.x) {
                const uint4 c = t4[i];
                cells_sum += (unsigned long long)c.x + (unsigned long long)c.y + (unsigned long long)c.z + (unsigned long long)c.w;
                float4 v;
                v.x = fmaf((float)c.x, inv_q, bg);
                v.y = fmaf((float)c.y, inv_q, bg);
                v.z = fmaf((float)c.z, inv_q, bg);
                v.w = fmaf((float)c.w, inv_q, bg);
                reinterpret_cast<float4*>(dst0)[i] = v;
            }
            flushed = true;
        } else {
            for (int i = threadIdx.x; i < n_tile; i += blockDim.x) cells_sum += (unsigned long long)tile_u[i];
        }
        // exact: the sum of the cells equals the sum of what was added unless a 32-bit cell wrapped (one reduction of the
        // difference instead of one per side)
        if (block_sum_ll((long long)cells_sum - mass, scratch) != 0) {
            // a 32-bit cell wrapped: redo this slab in float (border splats were already sent)
            fixed = false;
            flushed = false;
            __syncthreads();
            for (int i = threadIdx.x; i < n_tile; i += blockDim.x) tile_f[i] = 0.f;
            __syncthreads();
            tile_accumulate_with<float, N_IN>(tile_f, img, load4, pose, grid, p_begin, p_end, ys, ye, band_lo, band_hi, false);
            __syncthreads();
        }
    } else {
        tile_accumulate_with<float, N_IN>(tile_f, img, load4, pose, grid, p_begin, p_end, ys, ye, band_lo, band_hi, border);
        __syncthreads();
    }
    if (flushed) return;

    auto cell_value = [&](int i) -> float { return fixed ? (float)tile_u[i] * inv_q : tile_f[i]; };
    float* __restrict__ dst = img + (int64_t)ys * g0;
    if (tp.exclusive) {
        if ((n_tile & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
            for (int i = threadIdx.x; i < n_tile / 4; i += blockDim.x) {
                float4 v;
                v.x = cell_value(4 * i) + bg;
                v.y = cell_value(4 * i + 1) + bg;
                v.z = cell_value(4 * i + 2) + bg;
                v.w = cell_value(4 * i + 3) + bg;
                reinterpret_cast<float4*>(dst)[i] = v;
            }
        } else {
            for (int i = threadIdx.x; i < n_tile; i += blockDim.x) dst[i] = cell_value(i) + bg;
        }
    } else {
        for (int i = threadIdx.x; i < n_tile; i += blockDim.x) {
            const float v = cell_value(i);
            if (v != 0.f) red_add(dst + i, v);
        }
    }
}

}  // namespace dpr
