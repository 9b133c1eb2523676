// dpr_forward_fast.cuh - the Float32 2-d forward splat kernel tuned for sm_100a (included by dpr_forward.cu).
//
// Same decomposition as fwd_splat_tile2d_kernel (CTA = pose x slab x point split, slab accumulated in shared
// memory, one coalesced flush), with the inner loop rebuilt around what the ncu captures of that kernel showed
// (profiles/): it was bound by issue slots, not by memory.
//   * the next chunk's point is prefetched into registers while the current one is processed (a TMA-staged variant,
//     cp.async.bulk per 1024-point chunk behind a full/empty mbarrier pair, measured SLOWER - 3.03 ms vs 2.03 ms on
//     config 2 - because one shared stage forces the 32 warps into lock-step and exposes the copy latency, while a
//     second stage costs 12 rows of tile; TMA is used where it fits: dpr_pullback_tma.cuh stages whole pose images);
//   * the transform and the stencil run on Blackwell's packed FP32x2 pipe (FMUL2/FADD2: both output dimensions in
//     one instruction), still unfused and in the reference's operation order, so cell assignment is bit-identical
//     to the scalar path (dpr_common.cuh::stencil) - see the ptxas note at the stencil below;
//   * accumulation is fixed-point on the native 32-bit ATOMS.ADD (see dpr_forward.cu header comment); a 64-bit mass
//     checksum detects a wrapped cell exactly and the CTA then redoes the slab with float CAS atomics;
//   * the ~5% of lanes whose stencil leaves the slab (image border, slab edge) are not handled inline - that made
//     almost every warp execute the slow path - but compacted into a per-warp queue and processed 32 at a time.
#pragma once
#include "dpr_common.cuh"

namespace dpr {

constexpr int kChunk = 1024;          // points per chunk = threads per CTA = length of a culling run
constexpr int kQueueCap = 64;         // per-warp deferred-point queue (ints)

// Fixed-point scale.  A contribution v = w * out_weight * point_weight = w * pw' * A with pw' = point_weight * 2^-em in
// [0, 1] (2^em >= max(point_weight); power-of-two scaling is exact) and A = out_weight * 2^em is accumulated as the integer
// rint(w * pw' * Q), Q = rint(A * 2^(F - e)), 2^e >= A, so 2^(F-1) <= Q <= 2^F.  The integer is produced WITHOUT a
// conversion: multiplying by Q * 2^-149 (the subnormal whose bit pattern is the integer Q) lands the product in the
// subnormal range, where round-to-nearest leaves exactly rint(.) in the mantissa bits - __float_as_int of the product
// IS the fixed-point value.  The rounding of Q is undone exactly at the flush (inv_q = A / Q).
__device__ __forceinline__ int point_weight_exponent(const float* __restrict__ pw_stats) {
    if (!pw_stats) return 0;
    const float wmax = __ldg(pw_stats);
    int em = 0;
    if (wmax > 0.f && wmax < 3e38f) frexpf(wmax, &em);
    return em;
}
// shared-memory reduction without return value on a 32-bit shared address (ATOMS.ADD RZ): keeps the address
// arithmetic in 32 bits and out of the generic-pointer conversion the compiler otherwise repeats per iteration
__device__ __forceinline__ void red_shared_u32(uint32_t addr, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}


struct FastTileParams {
    int slabs, splits, rows, band_lo, band_hi, exclusive;
    int fixed_bits;
    int per_split;             // points per split, a multiple of kChunk
    const float* pw_stats;     // {max, min, mean} of point_weight or NULL
    const float* aabb;         // per-kChunk bounding boxes of the (spatially sorted) points, or NULL: no culling
    int pitch;                 // words per tile row: g0, or g0 + 4 (bank skew for spatially sorted points)
};

// shared-memory carve-up (bytes) after the tile
__host__ __device__ inline size_t fast_extra_smem(bool cull) {
    // per-warp deferred-point queues [+ surviving-chunk list when slab CTAs cull point runs]
    return (size_t)32 * kQueueCap * 4 + (cull ? (size_t)kChunk * 4 : 0);
}

// MODE 0: one slab per pose (whole image, hybrid band, or point splits); 1: several slabs; 2: several slabs with
// spatially sorted points and per-run culling.
template <int N_IN, bool HAS_PW, int MODE>
__global__ void __launch_bounds__(1024, 1)
fwd_tile2d_fast_kernel(const float* __restrict__ points, const float* __restrict__ rotation,
                       const float* __restrict__ translation, const float* __restrict__ background,
                       const float* __restrict__ out_weight, const float* __restrict__ point_weight,
                       float* __restrict__ out, Grid<float, 2> grid, int P, FastTileParams tp) {
    constexpr bool CULL = MODE == 2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ long long scratch[32];
    const int per_pose = tp.slabs * tp.splits;
    const int64_t b = blockIdx.x / per_pose;
    const int rem = blockIdx.x % per_pose;
    const int s = rem / tp.splits, q = rem % tp.splits;
    const int g0 = grid.g[0], g1 = grid.g[1];
    const int ys = tp.band_lo + s * tp.rows;
    const int ye = (ys + tp.rows < tp.band_hi) ? ys + tp.rows : tp.band_hi;
    const int nrows = ye - ys;
    const int pitch = tp.pitch;
    const int n_tile = nrows * pitch;                        // incl. the padding words of every row (they stay zero)
    const int tile_cap = tp.rows * pitch;                    // carve-up is the same for every CTA of the launch
    unsigned* tile_u = reinterpret_cast<unsigned*>(smem_raw);
    float* tile_f = reinterpret_cast<float*>(smem_raw);
    unsigned char* after = smem_raw + (((size_t)tile_cap * 4 + 127) / 128) * 128;
    int* queue = reinterpret_cast<int*>(after);
    int* clist = queue + 32 * kQueueCap;             // chunks of the current round that can touch this slab
    __shared__ int s_count, s_count_inner;
    float* __restrict__ img = out + b * grid.cells;
    const float bg = background ? __ldg(background + b) : 0.f;
    const bool border = (tp.band_lo > 0 || tp.band_hi < g1);
    const bool do_border = border && s == 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int* wq = queue + warp * kQueueCap;

    Pose<float, N_IN, 2> pose;
    load_pose(pose, rotation, translation, out_weight, b);
    const int p_begin = q * tp.per_split;
    const int p_end = (p_begin + tp.per_split < P) ? p_begin + tp.per_split : P;

    // fixed-point scale (uniform across the CTA, see point_weight_exponent above); not eligible -> float CAS accumulation
    bool fixed = false;
    float wq_den = 0.f, inv_qscale = 0.f;
    const int em = point_weight_exponent(HAS_PW ? tp.pw_stats : nullptr);
    const float pw_scale = ldexpf(1.0f, -em);
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
            inv_qscale = A / Q;
            fixed = true;
        }
    }

    for (int i = threadIdx.x; i < n_tile; i += blockDim.x) tile_u[i] = 0u;
    if (tp.exclusive && border) {
        const int lo_cells = tp.band_lo * g0;
        for (int i = threadIdx.x; i < lo_cells; i += blockDim.x) img[i] = bg;
        for (int i = tp.band_hi * g0 + threadIdx.x; i < g0 * g1; i += blockDim.x) img[i] = bg;
        __threadfence();
    }
    __syncthreads();

    long long mass = 0;
    if (fixed) {
        // ---- packed pose registers ------------------------------------------------------------------------
        float2 Rj[N_IN];
#pragma unroll
        for (int j = 0; j < N_IN; ++j) Rj[j] = make_float2(pose.R[0][j], pose.R[1][j]);
        const float2 neg_origin = make_float2(-pose.origin[0], -pose.origin[1]);   // proj - origin == proj + (-origin)
        const float2 scale2 = make_float2(grid.scale[0], grid.scale[1]);
        const float wq_pose = wq_den;
        const uint32_t tile_s = smem_u32(tile_u);           // 32-bit shared address: no generic-pointer conversion per add
        unsigned mass32 = 0;
        int wq_count = 0;                                   // warp-uniform

        // quantised add of one corner into the tile: the subnormal product's bit pattern is the integer
        auto tile_add = [&](int off, float v_times_q) {
            const uint32_t qv = __float_as_uint(v_times_q);
            red_shared_u32(tile_s + (uint32_t)off * 4u, qv);
            mass32 += qv;
        };
        // generic (any position) handling of one point: used for the deferred lanes
        auto slow_point = [&](int p) {
            float x[N_IN];
            load_point(x, points, (int64_t)p);
            int i0[2];
            float dl[2];
            if (!stencil(x, pose, grid, i0, dl)) return;
            const float pw = HAS_PW ? __ldg(point_weight + p) : 1.f;
            const float weight = pose.ow * pw;
            const float wqd = HAS_PW ? wq_den * (pw * pw_scale) : wq_den;
            const float du0 = 1.f - dl[0], du1 = 1.f - dl[1];
            const float w[4] = {du0 * du1, dl[0] * du1, du0 * dl[1], dl[0] * dl[1]};
            const bool x_lo = i0[0] >= 0, x_hi = i0[0] + 1 < g0;
#pragma unroll
            for (int cy = 0; cy < 2; ++cy) {
                const int iy = i0[1] + cy;
                if (iy < 0 || iy >= g1) continue;
                if (iy >= ys && iy < ye) {
                    const int off = (iy - ys) * pitch + i0[0];
                    if (x_lo) tile_add(off, w[2 * cy] * wqd);
                    if (x_hi) tile_add(off + 1, w[2 * cy + 1] * wqd);
                } else if (do_border && (iy < tp.band_lo || iy >= tp.band_hi)) {
                    float* addr = img + (int64_t)iy * g0 + i0[0];
                    const float va = w[2 * cy] * weight, vb = w[2 * cy + 1] * weight;
                    if (x_lo && x_hi) red_add2(addr, va, vb);
                    else if (x_lo) red_add(addr, va);
                    else if (x_hi) red_add(addr + 1, vb);
                }
            }
        };

        const int n_chunks = (p_end - p_begin + kChunk - 1) / kChunk;
        for (int round = 0; round < n_chunks; round += (CULL ? kChunk : n_chunks)) {
        // ---- which chunks of this round can touch the slab?  (spatially sorted points + per-chunk boxes) ----------
        int n_iter = CULL ? ((n_chunks - round < kChunk) ? n_chunks - round : kChunk) : n_chunks;
        int n_inner = 0;          // CULL: runs whose points are ALL interior to this slab (list grows from the front)
        if constexpr (CULL) {
            if (threadIdx.x == 0) { s_count = 0; s_count_inner = 0; }
            __syncthreads();
            const int c = round + (int)threadIdx.x;
            bool keep = false, inner = false;
            if (c < n_chunks) {
                const float* bx = tp.aabb + ((int64_t)(p_begin / kChunk) + c) * 2 * N_IN;
                float cx = -pose.origin[0], hx = 0.f, cy = -pose.origin[1], hy = 0.f;
#pragma unroll
                for (int j = 0; j < N_IN; ++j) {
                    const float bc = __ldg(bx + j), bh = __ldg(bx + N_IN + j);
                    cx = fmaf(pose.R[0][j], bc, cx);
                    hx = fmaf(fabsf(pose.R[0][j]), bh, hx);
                    cy = fmaf(pose.R[1][j], bc, cy);
                    hy = fmaf(fabsf(pose.R[1][j]), bh, hy);
                }
                cx *= grid.scale[0];
                cy *= grid.scale[1];
                hx = hx * grid.scale[0] + 0.05f + 1e-3f * fabsf(cx);     // extent of the projected box + rounding slack
                hy = hy * grid.scale[1] + 0.05f + 1e-3f * fabsf(cy);
                // a point at row coordinate y touches rows [y - 1.5, y + 0.5]; one more row of slack for the culling
                keep = !(cy + hy + 1.5f < (float)ys) && !(cy - hy - 2.5f > (float)(ye - 1));
                // every point of a full run has all four corners on chip iff the projected box satisfies
                // 0.5 < x <= g0 - 0.5 and ys + 0.5 < y <= ye - 0.5
                inner = keep && p_begin + (c + 1) * kChunk <= p_end && cx - hx > 0.5f && cx + hx < (float)g0 - 0.5f &&
                        cy - hy > (float)ys + 0.5f && cy + hy < (float)ye - 0.5f;
            }
            // two lists in the same array: inner runs from the front, the others from the back
            const bool mixed = keep && !inner;
            const unsigned mi = __ballot_sync(0xffffffffu, inner), mm = __ballot_sync(0xffffffffu, mixed);
            int bi = 0, bm = 0;
            if (lane == 0 && mi) bi = atomicAdd(&s_count_inner, __popc(mi));
            if (lane == 0 && mm) bm = atomicAdd(&s_count, __popc(mm));
            bi = __shfl_sync(0xffffffffu, bi, 0);
            bm = __shfl_sync(0xffffffffu, bm, 0);
            if (inner) clist[bi + __popc(mi & ((1u << lane) - 1u))] = c;
            if (mixed) clist[kChunk - 1 - (bm + __popc(mm & ((1u << lane) - 1u)))] = c;
            __syncthreads();
            n_iter = s_count;
            n_inner = s_count_inner;
        }
        if constexpr (CULL) {
            // ---- runs that are interior as a whole: no activity / bounds / slab test, no ballot, no queue ---------
            float xi[N_IN], pwi = 1.f;
            if (n_inner > 0) {
                const int p0 = p_begin + clist[0] * kChunk + (int)threadIdx.x;
                load_point(xi, points, (int64_t)p0);
                if (HAS_PW) pwi = __ldg(point_weight + p0);
            }
            for (int it0 = 0; it0 < n_inner; it0 += 64) {
                const int it1 = (it0 + 64 < n_inner) ? it0 + 64 : n_inner;
#pragma unroll 2
                for (int it = it0; it < it1; ++it) {
                    float x[N_IN];
#pragma unroll
                    for (int j = 0; j < N_IN; ++j) x[j] = xi[j];
                    const float pw = pwi;
                    if (it + 1 < n_inner) {
                        const int pn = p_begin + clist[it + 1] * kChunk + (int)threadIdx.x;
                        load_point(xi, points, (int64_t)pn);
                        if (HAS_PW) pwi = __ldg(point_weight + pn);
                    }
                    float2 prod[N_IN];
#pragma unroll
                    for (int j = 0; j < N_IN; ++j) prod[j] = __fmul2_rn(Rj[j], make_float2(x[j], x[j]));
                    float s0 = prod[0].x, s1 = prod[0].y;
#pragma unroll
                    for (int j = 1; j < N_IN; ++j) { s0 = __fadd_rn(s0, prod[j].x); s1 = __fadd_rn(s1, prod[j].y); }
                    const float2 coord = __fmul2_rn(__fadd2_rn(make_float2(s0, s1), neg_origin), scale2);
                    const float2 r = make_float2(ceilf(__fadd_rn(coord.x, -0.5f)), ceilf(__fadd_rn(coord.y, -0.5f)));
                    const float2 t = __fadd2_rn(r, make_float2(-0.5f, -0.5f));
                    const float2 dl = make_float2(__fsub_rn(coord.x, t.x), __fsub_rn(coord.y, t.y));
                    const float2 du = __fadd2_rn(make_float2(1.f, 1.f), make_float2(-dl.x, -dl.y));
                    const int ix = __float2int_rn(r.x) - 1, iy = __float2int_rn(r.y) - 1;
                    const float wq = HAS_PW ? wq_pose * (pw * pw_scale) : wq_pose;
                    const float a = du.y * wq, bq = dl.y * wq;
                    const int off = (iy - ys) * pitch + ix;
                    tile_add(off, du.x * a);
                    tile_add(off + 1, dl.x * a);
                    tile_add(off + pitch, du.x * bq);
                    tile_add(off + pitch + 1, dl.x * bq);
                }
                mass += mass32;
                mass32 = 0;
            }
        }
        // chunk of iteration i of the checked loop (with culling: the list that grows from the back)
        auto chunk_of = [&](int i) -> int { if constexpr (CULL) return clist[kChunk - 1 - i]; else return round + i; };
        // software pipeline: the point of the next chunk is loaded while the current one is processed
        float xn[N_IN], pwn = 1.f;
        if (n_iter > 0) {
            const int p0 = p_begin + chunk_of(0) * kChunk + (int)threadIdx.x;
            if (p0 < p_end) {
                load_point(xn, points, (int64_t)p0);
                if (HAS_PW) pwn = __ldg(point_weight + p0);
            }
        }
        for (int it0 = 0; it0 < n_iter; it0 += 64) {      // the 32-bit mass partial is folded into 64 bits every 64 chunks
        const int it1 = (it0 + 64 < n_iter) ? it0 + 64 : n_iter;
#pragma unroll 2
        for (int it = it0; it < it1; ++it) {
            const int c = chunk_of(it);
            const int c0 = p_begin + c * kChunk;
            const bool active = c0 + (int)threadIdx.x < p_end;
            float x[N_IN];
#pragma unroll
            for (int j = 0; j < N_IN; ++j) x[j] = xn[j];
            const float pw = pwn;
            if (it + 1 < n_iter) {
                const int pn = p_begin + chunk_of(it + 1) * kChunk + (int)threadIdx.x;
                if (pn < p_end) {
                    load_point(xn, points, (int64_t)pn);
                    if (HAS_PW) pwn = __ldg(point_weight + pn);
                }
            }

            // ---- transform + stencil, both output dimensions packed (src/raster.jl:88-99) -------------------
            // ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even though both carry .rn (and even with
            // -fmad=false), which would change the last bit of `coord` and flip cells.  Scalar add.rn.f32 IS respected,
            // so: products packed (FMUL2), every add that consumes a product scalar (FADD), the rest packed.
            // tests/test_abi.py asserts that this kernel contains no FFMA2 at all.
            float2 prod[N_IN];
#pragma unroll
            for (int j = 0; j < N_IN; ++j) prod[j] = __fmul2_rn(Rj[j], make_float2(x[j], x[j]));
            float s0 = prod[0].x, s1 = prod[0].y;
#pragma unroll
            for (int j = 1; j < N_IN; ++j) { s0 = __fadd_rn(s0, prod[j].x); s1 = __fadd_rn(s1, prod[j].y); }
            const float2 coord = __fmul2_rn(__fadd2_rn(make_float2(s0, s1), neg_origin), scale2);
            const float2 r = make_float2(ceilf(__fadd_rn(coord.x, -0.5f)), ceilf(__fadd_rn(coord.y, -0.5f)));
            const float2 t = __fadd2_rn(r, make_float2(-0.5f, -0.5f));
            const float2 dl = make_float2(__fsub_rn(coord.x, t.x), __fsub_rn(coord.y, t.y));   // coord - (r - 0.5)
            const float2 du = __fadd2_rn(make_float2(1.f, 1.f), make_float2(-dl.x, -dl.y));
            const int ix = __float2int_rn(r.x) - 1, iy = __float2int_rn(r.y) - 1;
            const int ry = iy - ys;
            const bool interior = active && (unsigned)ix < (unsigned)(g0 - 1) && (unsigned)ry < (unsigned)(nrows - 1);
            if (interior) {
                const float wq = HAS_PW ? wq_pose * (pw * pw_scale) : wq_pose;
                const float a = du.y * wq, bq = dl.y * wq;
                const int off = ry * pitch + ix;
                tile_add(off, du.x * a);
                tile_add(off + 1, dl.x * a);
                tile_add(off + pitch, du.x * bq);
                tile_add(off + pitch + 1, dl.x * bq);
            }
            // ---- defer the lanes that are not interior but touch this CTA's cells (warp-level compaction) ------
            bool slow = active && !interior;
            if constexpr (MODE != 0) {
                // several slabs: a point of another slab is not this CTA's business (there are no border rows then)
                slow = slow && (unsigned)(ry + 1) < (unsigned)(nrows + 1);
            }
            // MODE 0: every non-interior point either straddles the slab edge, lies in the border rows handled with
            // REDG by this CTA, or is (partly) outside the image: slow_point sorts that out
            const unsigned slow_mask = __ballot_sync(0xffffffffu, slow);
            if (slow_mask) {
                if (slow) wq[wq_count + __popc(slow_mask & ((1u << lane) - 1u))] = c0 + (int)threadIdx.x;
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
        if constexpr (CULL) __syncthreads();     // the chunk list is rebuilt in the next round
        }
        if (lane < wq_count) slow_point(wq[lane]);
        mass += mass32;

        mass = block_sum_ll(mass, scratch);          // contains the __syncthreads that ends the accumulation
        long long cells_sum = 0;
        for (int i = threadIdx.x; i < n_tile; i += blockDim.x) cells_sum += (long long)tile_u[i];
        cells_sum = block_sum_ll(cells_sum, scratch);
        if (cells_sum != mass) {
            // a 32-bit cell wrapped: redo this slab in float (border splats were already sent)
            fixed = false;
            for (int i = threadIdx.x; i < n_tile; i += blockDim.x) tile_f[i] = 0.f;
            __syncthreads();
            tile_accumulate<float, N_IN>(tile_f, img, points, point_weight, pose, grid, p_begin, p_end, ys, ye, tp.band_lo,
                                         tp.band_hi, false, pitch);
            __syncthreads();
        }
    } else {
        tile_accumulate<float, N_IN>(tile_f, img, points, point_weight, pose, grid, p_begin, p_end, ys, ye, tp.band_lo,
                                     tp.band_hi, do_border, pitch);
        __syncthreads();
    }

    // flush row by row (the tile pitch may differ from the image's): a warp per row, 16-byte accesses where aligned
    auto cell_value = [&](int i) -> float { return fixed ? (float)tile_u[i] * inv_qscale : tile_f[i]; };
    float* __restrict__ dst = img + (int64_t)ys * g0;
    const bool vec = (g0 & 3) == 0 && (pitch & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
    for (int r = warp; r < nrows; r += 32) {
        float* __restrict__ drow = dst + (int64_t)r * g0;
        const int t0 = r * pitch;
        if (tp.exclusive) {
            if (vec) {
                for (int c = lane; c < g0 / 4; c += 32) {
                    float4 v;
                    if (fixed) {
                        const uint4 q4 = reinterpret_cast<const uint4*>(tile_u + t0)[c];
                        v = make_float4(fmaf((float)q4.x, inv_qscale, bg), fmaf((float)q4.y, inv_qscale, bg),
                                        fmaf((float)q4.z, inv_qscale, bg), fmaf((float)q4.w, inv_qscale, bg));
                    } else {
                        const float4 f4 = reinterpret_cast<const float4*>(tile_f + t0)[c];
                        v = make_float4(f4.x + bg, f4.y + bg, f4.z + bg, f4.w + bg);
                    }
                    reinterpret_cast<float4*>(drow)[c] = v;
                }
            } else {
                for (int c = lane; c < g0; c += 32) drow[c] = cell_value(t0 + c) + bg;
            }
        } else {
            for (int c = lane; c < g0; c += 32) {
                const float v = cell_value(t0 + c);
                if (v != 0.f) red_add(drow + c, v);
            }
        }
    }
}

}  // namespace dpr
