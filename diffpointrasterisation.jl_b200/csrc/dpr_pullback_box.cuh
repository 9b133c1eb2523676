// dpr_pullback_box.cuh - 2-d Float32 pullback with the pose image's relevant BOX staged in shared memory, included by
// dpr_pullback.cu.
//
// pullback_gather2d_kernel (dpr_pullback_fast.cuh) gathers the 2 x 2 stencils straight from L1 / L2 and is bound by the
// L1 data pipe: the 32 lanes of a load land in a blob of the pose image ~20 pixels tall, i.e. in ~17 different 128-byte
// lines (profiles/ncu_full_r01_v9_cfg2_summary.csv: 16.8 tag requests per load instruction, data pipe 91 % busy).
// A CTA owns 1024 SPATIALLY SORTED points - a compact blob in space, hence a compact blob of pixels under every pose -
// so per (CTA, pose) the 64 x 64 pixel box around the projected centroid of the CTA's points is copied into shared
// memory with 16-byte cp.async (LDGSTS, two stages: the box of pose b+1 is in flight while pose b is gathered) and the
// stencils are read with LDS from a pitch-68 tile (bank = (x + 4 y) mod 32).  tools/probe_box.cu ->
// profiles/probe_box_r01.json: 1.6 - 1.7 T corner loads/s with the box copies included, against 0.35 - 0.69 T/s for direct
// __ldg gathers in the same geometry (the real kernel sustains 0.93 T/s).  No tensor map is involved (cp.async.bulk.tensor
// does not run on this pool); the copies are plain per-thread 16-byte asynchronous copies.
//
// Everything else - the point-owning decomposition, the stencil, the gradient arithmetic, the butterfly reduction, the
// d_background CTAs - is the code of pullback_gather2d_kernel, so the results are bit-identical to it: a warp whose 32
// stencils of a point slot all lie inside the box reads them from shared memory, any other warp takes the predicated
// global loads of the L1 kernel for that slot (the same values either way).
#pragma once
#include "dpr_pullback_fast.cuh"

namespace dpr {

constexpr int kBoxSize = 64;               // box edge in pixels
constexpr int kBoxPitch = kBoxSize + 4;    // floats per tile row: 16-byte aligned rows, bank = (x + 4 y) mod 32
constexpr int kBoxStageFloats = kBoxSize * kBoxPitch;

// dynamic shared memory of pullback_box2d_kernel: pose records + accumulators (as in pullback_gather2d_kernel), two box stages
inline size_t box_pullback_smem(int pose_chunk, int n_in) {
    const int NV = 2 * n_in + 3, PP = (NV + 3) / 4 * 4;
    size_t head = sizeof(float) * (size_t)pose_chunk * (PP + NV);
    if (head < sizeof(float) * (size_t)(pose_chunk * PP + 8)) head = sizeof(float) * (size_t)(pose_chunk * PP + 8);
    head = (head + 15) / 16 * 16;
    return head + sizeof(float) * 2 * kBoxStageFloats;
}

template <int N_IN, int K, bool HAS_PW>
__global__ void __launch_bounds__(256, 3)
pullback_box2d_kernel(const float* __restrict__ ds_dout, const float* __restrict__ points, const float* __restrict__ rotation,
                      const float* __restrict__ translation, const float* __restrict__ out_weight,
                      const float* __restrict__ point_weight, float* __restrict__ d_points, float* __restrict__ d_rotation,
                      float* __restrict__ d_translation, float* __restrict__ d_out_weight, float* __restrict__ d_point_weight,
                      const int32_t* __restrict__ perm, Grid<float, 2> grid, int P, int64_t B, int point_chunks,
                      int pose_chunk, float* __restrict__ d_background, int bg_ctas, int head_bytes) {
    using T = float;
    constexpr int NR = 2 * N_IN;            // rotation entries
    constexpr int NV = NR + 3;              // + translation (2) + out_weight
    constexpr int PP = (NV + 3) / 4 * 4;    // padded pose-parameter record: R (col-major), -origin (2), ow
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* pose_par = reinterpret_cast<T*>(smem_raw);              // [pose_chunk][PP]
    T* pose_acc = pose_par + (size_t)pose_chunk * PP;          // [pose_chunk][NV]
    T* stage0 = reinterpret_cast<T*>(smem_raw + head_bytes);   // [2][kBoxSize][kBoxPitch], 16-byte aligned

    // CTA layout as in pullback_gather2d_kernel: per pose chunk, point_chunks gather CTAs, then bg_ctas d_background CTAs
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
        T* warp_part = pose_acc;                               // 8 partial sums (the dynamic smem holds >= 8 values)
        for (int bl = lo; bl < hi; ++bl) {
            // images are 16-byte aligned with cells % 4 == 0 (checked by the host for this kernel)
            const float4* __restrict__ v4 = reinterpret_cast<const float4*>(ds_dout + (b0 + bl) * cells);
            T acc0 = T(0), acc1 = T(0);
            for (int64_t i = threadIdx.x; i < cells / 4; i += blockDim.x) {
                const float4 q = __ldg(v4 + i);
                acc0 += q.x + q.y;
                acc1 += q.z + q.w;
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

    T x[K][N_IN], pw[K], dpt[K][N_IN], dpw[K];
    bool valid[K];
    T csum[N_IN], cnt = T(0);
#pragma unroll
    for (int j = 0; j < N_IN; ++j) csum[j] = T(0);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int p = (pc * K + k) * (int)blockDim.x + (int)threadIdx.x;
        valid[k] = p < P;
        const int pp = valid[k] ? p : 0;
        load_point(x[k], points, (int64_t)pp);
        if (!valid[k]) {
            // padding lanes sit on the origin and never load (see pullback_gather2d_kernel)
#pragma unroll
            for (int j = 0; j < N_IN; ++j) x[k][j] = T(0);
        } else {
            cnt += T(1);
#pragma unroll
            for (int j = 0; j < N_IN; ++j) csum[j] += x[k][j];
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

    // centroid of the CTA's points (any finite point works: a bad centre only sends warps to the global-load path).
    // Scratch = the second box stage, which nothing writes before the first barrier of the pose loop.
    T ctr[N_IN];
    {
        T* scratch = stage0 + kBoxStageFloats;                 // [8 warps][N_IN + 1]
        cnt = warp_sum(cnt);
#pragma unroll
        for (int j = 0; j < N_IN; ++j) csum[j] = warp_sum(csum[j]);
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < N_IN; ++j) scratch[(threadIdx.x >> 5) * (N_IN + 1) + j] = csum[j];
            scratch[(threadIdx.x >> 5) * (N_IN + 1) + N_IN] = cnt;
        }
        __syncthreads();                                       // also publishes pose_par / pose_acc
        T n = T(0);
#pragma unroll
        for (int j = 0; j < N_IN; ++j) ctr[j] = T(0);
        for (int w = 0; w < 8; ++w) {
#pragma unroll
            for (int j = 0; j < N_IN; ++j) ctr[j] += scratch[w * (N_IN + 1) + j];
            n += scratch[w * (N_IN + 1) + N_IN];
        }
        const T inv = n > T(0) ? T(1) / n : T(0);
#pragma unroll
        for (int j = 0; j < N_IN; ++j) ctr[j] *= inv;
        __syncthreads();                                       // everyone has read the scratch before stage 1 is filled
    }

    // box origin for pose bl: the box is centred on the projected centroid, 16-byte aligned in x and clamped into the
    // image (g0, g1 >= kBoxSize and g0 % 4 == 0 are host-side conditions), so every in-box stencil is in bounds.
    // Every thread computes it from the same values, so it is uniform across the CTA.
    auto box_origin = [&](int bl, int& bx, int& by) {
        T par[PP];
#pragma unroll
        for (int i = 0; i < PP / 4; ++i) {
            const float4 q = reinterpret_cast<const float4*>(pose_par + bl * PP)[i];
            par[4 * i] = q.x; par[4 * i + 1] = q.y; par[4 * i + 2] = q.z; par[4 * i + 3] = q.w;
        }
        T c0 = par[NR], c1 = par[NR + 1];
#pragma unroll
        for (int j = 0; j < N_IN; ++j) { c0 = fmaf(par[2 * j], ctr[j], c0); c1 = fmaf(par[2 * j + 1], ctr[j], c1); }
        // clamp in floating point first (fmaxf / fminf drop a NaN), so the integer arithmetic below cannot overflow
        c0 = fminf(fmaxf(c0 * scale[0], -1.0e6f), 1.0e6f);
        c1 = fminf(fmaxf(c1 * scale[1], -1.0e6f), 1.0e6f);
        int ix = __float2int_rd(c0) - kBoxSize / 2, iy = __float2int_rd(c1) - kBoxSize / 2;
        ix &= ~3;
        bx = max(0, min(ix, g[0] - kBoxSize));
        by = max(0, min(iy, g[1] - kBoxSize));
    };
    // 64 rows x 16 chunks of 16 bytes = 1024 asynchronous copies, four per thread
    auto issue = [&](int bl, int bx, int by) {
        const T* src = ds_dout + (b0 + bl) * grid.cells + (int64_t)by * g[0] + bx + (threadIdx.x & 15) * 4;
        const uint32_t dst = smem_u32(stage0 + (bl & 1) * kBoxStageFloats + (threadIdx.x >> 4) * kBoxPitch + (threadIdx.x & 15) * 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = (threadIdx.x >> 4) + 16 * i;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)(16 * i * kBoxPitch * 4)),
                         "l"(src + (int64_t)row * g[0])
                         : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    int bx, by;
    box_origin(0, bx, by);
    issue(0, bx, by);
    for (int bl = 0; bl < n_pose; ++bl) {
        int nbx = 0, nby = 0;
        if (bl + 1 < n_pose) {
            box_origin(bl + 1, nbx, nby);
            issue(bl + 1, nbx, nby);                           // stage (bl+1)&1 was last read in iteration bl-1 (barrier below)
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();                                       // the box of pose bl is complete and visible
        const T* tile = stage0 + (bl & 1) * kBoxStageFloats;

        T par[PP];
#pragma unroll
        for (int i = 0; i < PP / 4; ++i) {
            const float4 q = reinterpret_cast<const float4*>(pose_par + bl * PP)[i];
            par[4 * i] = q.x; par[4 * i + 1] = q.y; par[4 * i + 2] = q.z; par[4 * i + 3] = q.w;
        }
        T R[2][N_IN];
#pragma unroll
        for (int j = 0; j < N_IN; ++j) { R[0][j] = par[2 * j]; R[1][j] = par[2 * j + 1]; }
        const T neg_origin[2] = {par[NR], par[NR + 1]};
        const T ow = par[NR + 2];
        const T ows[2] = {ow * scale[0], ow * scale[1]};
        const T* img = ds_dout + (b0 + bl) * grid.cells;
        asm volatile("" : "+l"(img));                          // one opaque 64-bit base (see pullback_gather2d_kernel)

        T acc[8], acc_ow = T(0);    // acc: d_rotation (col-major, NR values), d_translation (2) [, padding]
#pragma unroll
        for (int v = 0; v < 8; ++v) acc[v] = T(0);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            int ix, iy;
            T dl[2];
            stencil2<T, N_IN>(x[k], R, neg_origin, scale, g, ix, iy, dl);
            // whole 2 x 2 stencil inside the box (hence inside the image)?  unsigned differences: no overflow for far-away points
            const unsigned rx = (unsigned)ix - (unsigned)bx, ry = (unsigned)iy - (unsigned)by;
            const bool in_box = rx < (unsigned)(kBoxSize - 1) && ry < (unsigned)(kBoxSize - 1);
            T G00 = T(0), G10 = T(0), G01 = T(0), G11 = T(0);
            if (__all_sync(0xffffffffu, in_box || !valid[k])) {
                if (valid[k]) {
                    const T* q = tile + ry * kBoxPitch + rx;
                    G00 = q[0]; G10 = q[1]; G01 = q[kBoxPitch]; G11 = q[kBoxPitch + 1];
                }
            } else {
                // per-corner bounds rule (src/raster_pullback.jl:51) as four load predicates
                const bool x_lo = valid[k] && (unsigned)ix < (unsigned)g[0], x_hi = valid[k] && (unsigned)(ix + 1) < (unsigned)g[0];
                const bool y_lo = (unsigned)iy < (unsigned)g[1], y_hi = (unsigned)(iy + 1) < (unsigned)g[1];
                const int off = iy * g[0] + ix;          // 32-bit: the host guarantees g0*g1 < 2^30; OOB lanes never load
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
            const T sx = gx * (HAS_PW ? ows[0] * pw[k] : ows[0]), sy = gy * (HAS_PW ? ows[1] * pw[k] : ows[1]);   // :60, :67
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
            int slot = vsel;
            T val = acc[0];
            bool mine = (lane & 3) == 0 && vsel < NV;
            if constexpr (N_IN == 3) {
                if (lane == 1) { slot = NV - 1; val = acc_ow; mine = true; }
            }
            if (mine) atomicAdd(&pose_acc[bl * NV + slot], val);
        }
        bx = nbx; by = nby;
        __syncthreads();                                       // stage bl&1 is refilled by the copy issued in iteration bl+1
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
    // (the trailing barrier of the pose loop has already ordered the pose_acc atomics before these reads)
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
