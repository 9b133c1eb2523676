// dpr_pullback_fast.cuh - 2-d pullback kernel (gathers from global memory / L2), included by dpr_pullback.cu.
//
// Same decomposition as pullback_gather_global_kernel (a thread owns K points and loops over its CTA's pose chunk),
// rebuilt after the first ncu capture (profiles/ncu_full_r01_v1_summary.csv: 194 instructions per warp-splat, LSU
// data pipe 80 % busy, issue slots 55 %):
//   * pose parameters of the chunk are staged once in shared memory and read back with broadcast LDS.128;
//   * the stencil runs on the packed FP32x2 pipe where that is safe (see the ptxas note in dpr_forward_fast.cuh);
//   * the four corners are fetched with predicated loads from one base address - no per-corner branches;
//   * the per-pose reduction of the 2*N_in+3 point-sums uses a transposing butterfly (each shuffle step halves the
//     number of live values) instead of one 5-step shuffle tree per value.
#pragma once
#include <type_traits>

#include "dpr_common.cuh"

namespace dpr {

// Sum 8 values across the warp.  On return lane L (with (L & 3) == 0) holds the total of value index
// v = 4*bit4(L) + 2*bit3(L) + bit2(L) in a[0].
template <typename T>
__device__ __forceinline__ void butterfly8(T (&a)[8], int lane) {
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const T send = h16 ? a[i] : a[i + 4];
        const T keep = h16 ? a[i + 4] : a[i];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const T send = h8 ? a[i] : a[i + 2];
        const T keep = h8 ? a[i + 2] : a[i];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    {
        const T send = h4 ? a[0] : a[1];
        const T keep = h4 ? a[1] : a[0];
        a[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    a[0] += __shfl_xor_sync(0xffffffffu, a[0], 2);
    a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1);
}

// Bilinear sample s = sum_c W_c G_c and its derivatives with respect to (dl0, dl1) from the four corner values, with the
// common subexpression c = G11 - G01 - G10 + G00 shared: gx = a0 + dl1 c, gy = b0 + dl0 c, s = G00 + dl0 a0 + dl1 gy
// (a0 = G10 - G00, b0 = G01 - G00).  8 instructions instead of 16 for the three separate formulas; differences of the
// corner values first, as before, so nothing cancels against the oracle's Float64 sums.
template <typename T>
__device__ __forceinline__ void bilinear_with_gradient(T G00, T G10, T G01, T G11, T dl0, T dl1, T& s, T& gx, T& gy) {
    const T a0 = G10 - G00, b0 = G01 - G00, c = (G11 - G01) - a0;
    gx = fma(dl1, c, a0);
    gy = fma(dl0, c, b0);
    s = fma(dl1, gy, fma(dl0, a0, G00));
}

// lower-corner cell (0-based) and dl for the 2-d case, bit-identical to dpr_common.cuh::stencil.
template <typename T, int N_IN>
__device__ __forceinline__ void stencil2(const T (&x)[N_IN], const T (&R)[2][N_IN], const T (&neg_origin)[2],
                                         const T (&scale)[2], const int (&g)[2], int& ix, int& iy, T (&dl)[2]) {
    T r[2];
    if constexpr (std::is_same<T, float>::value) {
        float2 prod[N_IN];
#pragma unroll
        for (int j = 0; j < N_IN; ++j) prod[j] = __fmul2_rn(make_float2(R[0][j], R[1][j]), make_float2(x[j], x[j]));
        float s0 = prod[0].x, s1 = prod[0].y;
#pragma unroll
        for (int j = 1; j < N_IN; ++j) { s0 = __fadd_rn(s0, prod[j].x); s1 = __fadd_rn(s1, prod[j].y); }
        const float2 coord = __fmul2_rn(__fadd2_rn(make_float2(s0, s1), make_float2(neg_origin[0], neg_origin[1])),
                                        make_float2(scale[0], scale[1]));
        r[0] = ceilf(__fadd_rn(coord.x, -0.5f));
        r[1] = ceilf(__fadd_rn(coord.y, -0.5f));
        const float2 t = __fadd2_rn(make_float2(r[0], r[1]), make_float2(-0.5f, -0.5f));
        dl[0] = __fsub_rn(coord.x, t.x);
        dl[1] = __fsub_rn(coord.y, t.y);
    } else {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            T proj = mul_rn(R[k][0], x[0]);
#pragma unroll
            for (int j = 1; j < N_IN; ++j) proj = add_rn(proj, mul_rn(R[k][j], x[j]));
            const T coord = mul_rn(add_rn(proj, neg_origin[k]), scale[k]);
            r[k] = ceil_t(sub_rn(coord, T(0.5)));
            dl[k] = sub_rn(coord, sub_rn(r[k], T(0.5)));
        }
    }
    // cvt.rni.s32 saturates (huge -> INT_MAX / INT_MIN, NaN -> 0), and the -1 is done modulo 2^32, so far-away points
    // end up with indices for which every corner predicate of the callers is false; a NaN coordinate maps to
    // index -1 (one in-bounds column, memory-safe) and its NaN weights propagate to the results.
    (void)g;
    ix = (int)((unsigned)to_int_sat(r[0]) - 1u);
    iy = (int)((unsigned)to_int_sat(r[1]) - 1u);
}

template <typename T, int N_IN, int K, bool HAS_PW, bool PAIR>
__global__ void __launch_bounds__(256)
pullback_gather2d_kernel(const T* __restrict__ ds_dout, const T* __restrict__ points, const T* __restrict__ rotation,
                         const T* __restrict__ translation, const T* __restrict__ out_weight,
                         const T* __restrict__ point_weight, T* __restrict__ d_points, T* __restrict__ d_rotation,
                         T* __restrict__ d_translation, T* __restrict__ d_out_weight, T* __restrict__ d_point_weight,
                         const int32_t* __restrict__ perm, Grid<T, 2> grid, int P, int64_t B, int point_chunks,
                         int pose_chunk, T* __restrict__ d_background, int bg_ctas) {
    constexpr int NR = 2 * N_IN;            // rotation entries
    constexpr int NV = NR + 3;              // + translation (2) + out_weight
    constexpr int PP = (NV + 3) / 4 * 4;    // padded pose-parameter record: R (col-major), -origin (2), ow
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* pose_par = reinterpret_cast<T*>(smem_raw);              // [pose_chunk][PP]
    T* pose_acc = pose_par + (size_t)pose_chunk * PP;          // [pose_chunk][NV]

    // CTA layout: per pose chunk, point_chunks gather CTAs followed by bg_ctas CTAs that compute d_background of the
    // chunk's poses (src/raster_pullback.jl:78).  CTAs are dispatched in blockIdx order, so those run while the chunk's
    // gather CTAs are resident and have pulled the chunk's images into L2, and they use HBM / L2 bandwidth the gather
    // CTAs (bound by the L1 data pipe) leave idle: the separate 0.17 ms background_sum pass of config 2 disappears.
    const int per_chunk = point_chunks + bg_ctas;
    const int pc = blockIdx.x % per_chunk;
    const int64_t bc = blockIdx.x / per_chunk;
    const int64_t b0 = bc * pose_chunk;
    const int n_pose = (int)((b0 + pose_chunk < B ? b0 + pose_chunk : B) - b0);
    if (pc >= point_chunks) {
        const int e = pc - point_chunks;
        const int per = (n_pose + bg_ctas - 1) / bg_ctas;
        const int lo = e * per, hi = (lo + per < n_pose) ? lo + per : n_pose;
        const int64_t cells = grid.cells;
        const bool vec = sizeof(T) == 4 && (cells & 3) == 0 && (reinterpret_cast<uintptr_t>(ds_dout) & 15) == 0;
        T* warp_part = pose_acc;                               // 8 partial sums (the dynamic smem holds >= 8 values)
        for (int bl = lo; bl < hi; ++bl) {
            const T* __restrict__ src = ds_dout + (b0 + bl) * cells;
            T acc0 = T(0), acc1 = T(0);
            if (vec) {
                const float4* __restrict__ v4 = reinterpret_cast<const float4*>(src);
                for (int64_t i = threadIdx.x; i < cells / 4; i += blockDim.x) {
                    const float4 q = __ldg(v4 + i);
                    acc0 += T(q.x) + T(q.y);
                    acc1 += T(q.z) + T(q.w);
                }
            } else {
                for (int64_t i = threadIdx.x; i < cells; i += blockDim.x) acc0 += __ldg(src + i);
            }
            T t = warp_sum(acc0 + acc1);
            if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = t;
            __syncthreads();
            if (threadIdx.x < 32) {
                t = threadIdx.x < (blockDim.x >> 5) ? warp_part[threadIdx.x] : T(0);
                t = warp_sum(t);
                if (threadIdx.x == 0) d_background[b0 + bl] = t;
            }
            __syncthreads();
        }
        return;
    }
    for (int i = threadIdx.x; i < n_pose * PP; i += blockDim.x) {
        const int bl = i / PP, v = i % PP;
        const int64_t b = b0 + bl;
        T val = T(0);
        if (v < NR) val = __ldg(rotation + b * NR + v);
        else if (v < NR + 2) val = -sub_rn(T(-1), __ldg(translation + b * 2 + (v - NR)));   // -origin, origin = -1 - t
        else if (v == NR + 2) val = out_weight ? __ldg(out_weight + b) : T(1);
        pose_par[i] = val;
    }
    for (int i = threadIdx.x; i < n_pose * NV; i += blockDim.x) pose_acc[i] = T(0);
    __syncthreads();

    T x[K][N_IN], pw[K], dpt[K][N_IN], dpw[K];
    bool valid[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int p = (pc * K + k) * (int)blockDim.x + (int)threadIdx.x;
        valid[k] = p < P;
        const int pp = valid[k] ? p : 0;
        load_point(x[k], points, (int64_t)pp);
        if (!valid[k]) {
            // padding lanes sit on the origin (finite weights for every pose) and never load: all four corner
            // predicates below include valid[k], so every contribution is an exact zero.  (A far-away point is NOT
            // safe: a matrix row orthogonal to it - e.g. (1,-1,0)/sqrt(2) against (c,c,c) - projects it to the origin.)
#pragma unroll
            for (int j = 0; j < N_IN; ++j) x[k][j] = T(0);
        }
        pw[k] = HAS_PW ? __ldg(point_weight + pp) : T(1);
#pragma unroll
        for (int j = 0; j < N_IN; ++j) dpt[k][j] = T(0);
        dpw[k] = T(0);
    }
    const int g[2] = {grid.g[0], grid.g[1]};
    const T scale[2] = {grid.scale[0], grid.scale[1]};
    const int lane = threadIdx.x & 31;
    const int vsel = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);

    for (int bl = 0; bl < n_pose; ++bl) {
        T par[PP];
        if constexpr (std::is_same<T, float>::value) {
#pragma unroll
            for (int i = 0; i < PP / 4; ++i) {
                const float4 q = reinterpret_cast<const float4*>(pose_par + bl * PP)[i];
                par[4 * i] = q.x; par[4 * i + 1] = q.y; par[4 * i + 2] = q.z; par[4 * i + 3] = q.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < PP / 2; ++i) {
                const double2 q = reinterpret_cast<const double2*>(pose_par + bl * PP)[i];
                par[2 * i] = q.x; par[2 * i + 1] = q.y;
            }
        }
        T R[2][N_IN];
#pragma unroll
        for (int j = 0; j < N_IN; ++j) { R[0][j] = par[2 * j]; R[1][j] = par[2 * j + 1]; }
        const T neg_origin[2] = {par[NR], par[NR + 1]};
        const T ow = par[NR + 2];
        const T ows[2] = {ow * scale[0], ow * scale[1]};
        const T* img = ds_dout + (b0 + bl) * grid.cells;
        // keep the pose image base as ONE opaque 64-bit register pair: otherwise the compiler re-derives
        // b*cells + offset with a 64-bit multiply-add and two LEAs for every corner load
        asm volatile("" : "+l"(img));

        T acc[8], acc_ow = T(0);    // acc: d_rotation (col-major, NR values), d_translation (2) [, padding]
#pragma unroll
        for (int v = 0; v < 8; ++v) acc[v] = T(0);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            int ix, iy;
            T dl[2];
            stencil2<T, N_IN>(x[k], R, neg_origin, scale, g, ix, iy, dl);
            // per-corner bounds rule (src/raster_pullback.jl:51) as four load predicates
            const bool x_lo = valid[k] && (unsigned)ix < (unsigned)g[0], x_hi = valid[k] && (unsigned)(ix + 1) < (unsigned)g[0];
            const bool y_lo = (unsigned)iy < (unsigned)g[1], y_hi = (unsigned)(iy + 1) < (unsigned)g[1];
            const int off = iy * g[0] + ix;          // 32-bit: the host guarantees g0*g1 < 2^30; OOB lanes never load
            T G00 = T(0), G10 = T(0), G01 = T(0), G11 = T(0);
            if constexpr (std::is_same<T, float>::value && PAIR) {
                // The L1 cost of a gather is per lane and instruction, and (ix, ix+1) share a 32-byte sector 7 times
                // out of 8: fetch the 8-byte aligned pair that holds ix with one LDG.64 and only the lanes with odd
                // ix issue a second 4-byte load (rows are 8-byte aligned: g0 even, checked by the host).
                const bool odd = ix & 1;
                const bool both = x_lo && x_hi;
                const float* pa = img + (off & ~1);          // one 64-bit address; everything else is an immediate/row offset
                const float* pr = pa + g[0];
                float2 q0 = make_float2(0.f, 0.f), q1 = make_float2(0.f, 0.f);
                float e0 = 0.f, e1 = 0.f;
                if (both && y_lo) q0 = __ldg(reinterpret_cast<const float2*>(pa));
                if (both && y_hi) q1 = __ldg(reinterpret_cast<const float2*>(pr));
                if (both && odd && y_lo) e0 = __ldg(pa + 2);
                if (both && odd && y_hi) e1 = __ldg(pr + 2);
                G00 = odd ? q0.y : q0.x; G10 = odd ? e0 : q0.y;
                G01 = odd ? q1.y : q1.x; G11 = odd ? e1 : q1.y;
                if (!both && (x_lo || x_hi)) {   // left / right image edge: one column only (rare)
                    const float* base = img + off;
                    if (x_lo && y_lo) G00 = __ldg(base);
                    if (x_hi && y_lo) G10 = __ldg(base + 1);
                    if (x_lo && y_hi) G01 = __ldg(base + g[0]);
                    if (x_hi && y_hi) G11 = __ldg(base + g[0] + 1);
                }
            } else {
                const T* base = img + off;
                if (x_lo && y_lo) G00 = __ldg(base);
                if (x_hi && y_lo) G10 = __ldg(base + 1);
                if (x_lo && y_hi) G01 = __ldg(base + g[0]);
                if (x_hi && y_hi) G11 = __ldg(base + g[0] + 1);
            }
            T s, gx, gy;     // s = sum_c W_c G_c; gx, gy = its derivatives with respect to the cell coordinate
            bilinear_with_gradient(G00, G10, G01, G11, dl[0], dl[1], s, gx, gy);
            acc_ow += HAS_PW ? s * pw[k] : s;                 // src/raster_pullback.jl:57
            dpw[k] += s * ow;                                  // :58
            // factor * scale (:60, :67) with the pose's part hoisted out of the point loop
            const T sx = gx * (HAS_PW ? ows[0] * pw[k] : ows[0]), sy = gy * (HAS_PW ? ows[1] * pw[k] : ows[1]);
            if constexpr (N_IN == 3) {
                acc[6] += sx; acc[7] += sy;                    // :68
            } else {
                acc[4] += sx; acc[5] += sy;
            }
#pragma unroll
            for (int j = 0; j < N_IN; ++j) {
                acc[2 * j] += sx * x[k][j];                    // :69
                acc[2 * j + 1] += sy * x[k][j];
                dpt[k][j] = fma(R[1][j], sy, fma(R[0][j], sx, dpt[k][j]));      // :70-71
            }
        }
        if constexpr (N_IN == 2) acc[6] = acc_ow;              // 7 values fit the butterfly
        butterfly8(acc, lane);
        if constexpr (N_IN == 3) acc_ow = warp_sum(acc_ow);
        {
            // butterfly slot -> (rotation entries | translation | out_weight) slot of pose_acc, ONE shared-memory
            // atomicAdd instruction (a CAS loop on sm_100a) for all of them: lanes 0, 4, .. 28 hold slot vsel
            // (N_IN==3: 0..5 rotation, 6..7 translation; N_IN==2: 0..3 rot, 4..5 trans, 6 ow), lane 1 the 9th value
            int slot = vsel;
            T val = acc[0];
            bool mine = (lane & 3) == 0 && vsel < NV;
            if constexpr (N_IN == 3) {
                if (lane == 1) { slot = NV - 1; val = acc_ow; mine = true; }
            }
            if (mine) atomicAdd(&pose_acc[bl * NV + slot], val);
        }
    }

#pragma unroll
    for (int k = 0; k < K; ++k) {
        if (!valid[k]) continue;
        int p = (pc * K + k) * (int)blockDim.x + (int)threadIdx.x;
        if (perm) p = __ldg(perm + p);       // points were spatially sorted: write through the permutation
#pragma unroll
        for (int j = 0; j < N_IN; ++j) red_add(d_points + (int64_t)p * N_IN + j, dpt[k][j]);
        if (d_point_weight) red_add(d_point_weight + p, dpw[k]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_pose * NV; i += blockDim.x) {
        const int bl = i / NV, v = i % NV;
        const T r = pose_acc[i];
        const int64_t b = b0 + bl;
        if (v < NR) red_add(d_rotation + b * NR + v, r);
        else if (v < NR + 2) red_add(d_translation + b * 2 + (v - NR), r);
        else if (d_out_weight) red_add(d_out_weight + b, r);
    }
}

}  // namespace dpr
