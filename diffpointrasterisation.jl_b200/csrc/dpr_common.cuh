// dpr_common.cuh - shared device helpers for the sm_100a splat / pullback kernels.
//
// The stencil here is the single place where a point is assigned to a cell.  It reproduces, operation by
// operation and with FMA contraction forbidden, the reference's reference_coordinate_and_deltas
// (/root/reference src/raster.jl:85-101) with origin = -1 - t (src/raster.jl:53): the gradient with respect to the
// coordinate jumps at cell centres, so a 1-ulp difference in `coord` would flip cells.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dpr {

// ---- exactly-rounded scalar ops (never contracted into FMA by nvcc) ------------------------------------
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float ceil_t(float a) { return ceilf(a); }
__device__ __forceinline__ double ceil_t(double a) { return ceil(a); }
__device__ __forceinline__ int to_int_sat(float a) { return __float2int_rn(a); }
__device__ __forceinline__ int to_int_sat(double a) { return __double2int_rn(a); }

// Pose parameters held in registers: R[k][j] = rotation[k + j*N_OUT] (column-major SMatrix),
// origin = -1 - t (src/raster.jl:53), ow = out_weight.
template <typename T, int N_IN, int N_OUT>
struct Pose {
    T R[N_OUT][N_IN];
    T origin[N_OUT];
    T ow;
};

template <typename T, int N_IN, int N_OUT>
__device__ __forceinline__ void load_pose(Pose<T, N_IN, N_OUT>& p, const T* __restrict__ rotation,
                                          const T* __restrict__ translation, const T* __restrict__ out_weight,
                                          int64_t b) {
#pragma unroll
    for (int j = 0; j < N_IN; ++j)
#pragma unroll
        for (int k = 0; k < N_OUT; ++k) p.R[k][j] = __ldg(rotation + b * (N_OUT * N_IN) + k + j * N_OUT);
#pragma unroll
    for (int k = 0; k < N_OUT; ++k) p.origin[k] = sub_rn(T(-1), __ldg(translation + b * N_OUT + k));
    p.ow = out_weight ? __ldg(out_weight + b) : T(1);
}

// Grid geometry: extents (int), scale = g/2 (src/raster.jl:25), strides in elements.
template <typename T, int N_OUT>
struct Grid {
    int g[N_OUT];
    T scale[N_OUT];
    int64_t cells;
};

// Lower-corner cell (0-based: i0 = ref - 1) and dl = distance to the lower cell centre.
// Returns false when no corner can be in bounds (or the coordinate is NaN); i0/dl are then unspecified.
//   proj  = ((R_k1 x_1 + R_k2 x_2) + R_k3 x_3)            src/raster.jl:88   (left to right, no FMA)
//   coord = (proj - origin) * scale                        src/raster.jl:92
//   ref   = ceil(coord - 0.5)                              src/raster.jl:94
//   dl    = coord - (ref - 0.5)                            src/raster.jl:97
template <typename T, int N_IN, int N_OUT>
__device__ __forceinline__ bool stencil(const T (&x)[N_IN], const Pose<T, N_IN, N_OUT>& pose,
                                        const Grid<T, N_OUT>& grid, int (&i0)[N_OUT], T (&dl)[N_OUT]) {
    bool any = true;
#pragma unroll
    for (int k = 0; k < N_OUT; ++k) {
        T proj = mul_rn(pose.R[k][0], x[0]);
#pragma unroll
        for (int j = 1; j < N_IN; ++j) proj = add_rn(proj, mul_rn(pose.R[k][j], x[j]));
        const T coord = mul_rn(sub_rn(proj, pose.origin[k]), grid.scale[k]);
        const T r = ceil_t(sub_rn(coord, T(0.5)));
        dl[k] = sub_rn(coord, sub_rn(r, T(0.5)));
        // 1-based ref must satisfy 0 <= ref <= g for at least one of {ref, ref+1} to lie in 1..g
        any = any && (r >= T(0)) && (r <= T(grid.g[k]));
        i0[k] = to_int_sat(r) - 1;
    }
    return any;
}

// Multilinear weight of corner c (bit k of c = shift in dimension k; dimension 0 is the least-significant bit,
// src/util.jl:7-8,26-27): shift 0 -> 1-dl, shift 1 -> dl (src/raster.jl:104-106), product left to right.
template <typename T, int N_OUT>
__device__ __forceinline__ T corner_weight(int c, const T (&dl)[N_OUT], const T (&du)[N_OUT]) {
    T w = (c & 1) ? dl[0] : du[0];
#pragma unroll
    for (int k = 1; k < N_OUT; ++k) w = w * (((c >> k) & 1) ? dl[k] : du[k]);
    return w;
}

template <typename T, int N_IN>
__device__ __forceinline__ void load_point(T (&x)[N_IN], const T* __restrict__ points, int64_t p) {
#pragma unroll
    for (int j = 0; j < N_IN; ++j) x[j] = __ldg(points + p * N_IN + j);
}

// ---- global reductions without return value (REDG) ------------------------------------------------------
__device__ __forceinline__ void red_add(float* addr, float v) { atomicAdd(addr, v); }
__device__ __forceinline__ void red_add(double* addr, double v) { atomicAdd(addr, v); }
// two adjacent elements; uses the native REDG.ADD.F32x2 when the pair is 8-byte aligned
__device__ __forceinline__ void red_add2(float* addr, float a, float b) {
    if ((reinterpret_cast<uintptr_t>(addr) & 7u) == 0) {
        atomicAdd(reinterpret_cast<float2*>(addr), make_float2(a, b));
    } else {
        atomicAdd(addr, a);
        atomicAdd(addr + 1, b);
    }
}
__device__ __forceinline__ void red_add2(double* addr, double a, double b) {
    atomicAdd(addr, a);
    atomicAdd(addr + 1, b);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- mbarrier + TMA (cp.async.bulk) helpers --------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-d bulk copy global -> shared through the TMA unit, completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

}  // namespace dpr
