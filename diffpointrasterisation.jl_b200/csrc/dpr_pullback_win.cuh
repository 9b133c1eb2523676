// dpr_pullback_win.cuh - 2-d Float32 pullback for pose images LARGER than shared memory: TMA-staged WINDOWS
// (included by dpr_pullback.cu).
//
// pullback_gather2d_kernel gathers ds_dout through L1 and is bound by the L1 data pipe (ncu: 87 % busy, 42 wavefronts
// per warp-splat) although its points are spatially sorted.  pullback_tma2d_kernel avoids that pipe cost by staging the
// whole pose image in shared memory, but a 256 x 256 Float32 image (config 2) does not fit.  It does not have to: a
// CTA owns a run of consecutive Morton-sorted points, a compact blob in space, and a pose maps a compact blob to a
// compact blob of pixels.  So per (CTA, pose) the producer warp
//   * projects the CENTROID of the CTA's points with the pose and places a band of `wy` full-width image rows around
//     it (clamped into the image).  Full-width rows are contiguous in memory, so the band is
//   * ONE cp.async.bulk (TMA, SASS UBLKCP) of wy * g0 * 4 bytes into the next stage of a shared-memory ring, completion
//     counted in bytes on an mbarrier,
//   * and publishes the band's first row next to the stage.
// The consumer warps take the four corners from the staged band when the stencil lies inside it (and inside the
// image horizontally) and from global memory with the per-corner bounds rule otherwise - through ONE branch-free
// sequence of predicated generic loads whose base pointer is selected per point.  Everything else - a thread owns K
// points across the pose chunk, transposing butterfly for the per-pose sums, REDG once per point and chunk - is the
// scheme of the other pullback kernels.  d_background is produced by background_sum_kernel (bands do not cover the
// image).
//
// Measured on the way here (config 2, B200): a 136 x 128 box per window copied row by row (128 bulk copies per
// window, issued by the 32 producer lanes) ran at 3.8 ms - the TMA unit's per-copy cost, not its bandwidth, was the
// limit - against 2.13 ms for the L1 gathers.  The natural single-instruction form, a 3-d tensor-map copy
// (cp.async.bulk.tensor / UTMALDG, zero fill outside the image), cannot be used on this pool's B200s: every UTMALDG,
// including the CUDA programming guide's own libcu++ example (tools/probe_tma_box2.cu), dies with "illegal
// instruction" although cuTensorMapEncodeTiled succeeds.
#pragma once
#include "dpr_common.cuh"
#include "dpr_pullback_fast.cuh"  // stencil2, butterfly8
#include "dpr_pullback_tma.cuh"   // kTmaConsumers, kTmaRound, named_bar_sync

namespace dpr {

// mean of every run of `chunk` consecutive (sorted) points; one CTA of 256 threads per run
template <int N_IN>
__global__ void __launch_bounds__(256) chunk_centroid_kernel(const float* __restrict__ points, int64_t P, int chunk,
                                                             float* __restrict__ centroid) {
    const int64_t lo = (int64_t)blockIdx.x * chunk;
    const int64_t hi = (lo + chunk < P) ? lo + chunk : P;
    float s[N_IN];
#pragma unroll
    for (int j = 0; j < N_IN; ++j) s[j] = 0.f;
    for (int64_t p = lo + threadIdx.x; p < hi; p += blockDim.x) {
#pragma unroll
        for (int j = 0; j < N_IN; ++j) {
            const float v = __ldg(points + p * N_IN + j);
            s[j] += (fabsf(v) < 1e18f) ? v : 0.f;       // non-finite / absurd points must not drag the window away
        }
    }
    __shared__ float part[8][N_IN];
#pragma unroll
    for (int j = 0; j < N_IN; ++j) {
        s[j] = warp_sum(s[j]);
        if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5][j] = s[j];
    }
    __syncthreads();
    if (threadIdx.x < N_IN) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += part[w][threadIdx.x];
        centroid[(int64_t)blockIdx.x * N_IN + threadIdx.x] = t / (float)(hi > lo ? hi - lo : 1);
    }
}

template <int N_IN, int K, bool HAS_PW, int STAGES>
__global__ void __launch_bounds__(kTmaConsumers + 32, 1)
pullback_win2d_kernel(const float* __restrict__ ds_dout,
                      const float* __restrict__ points, const float* __restrict__ centroid,
                      const float* __restrict__ rotation, const float* __restrict__ translation,
                      const float* __restrict__ out_weight, const float* __restrict__ point_weight,
                      float* __restrict__ d_points, float* __restrict__ d_rotation, float* __restrict__ d_translation,
                      float* __restrict__ d_out_weight, float* __restrict__ d_point_weight,
                      const int32_t* __restrict__ perm, Grid<float, 2> grid, int P, int64_t B, int point_chunks,
                      int pose_chunk, int wy) {
    constexpr int NR = 2 * N_IN, NV = NR + 3, PP = (NV + 3) / 4 * 4;
    const uint32_t band_bytes = (uint32_t)wy * (uint32_t)grid.g[0] * 4u;          // a multiple of 16 (g0 % 4 == 0)
    const size_t kBoxBytes = ((size_t)band_bytes + 127) / 128 * 128;                // stage stride
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* tiles = reinterpret_cast<float*>(smem_raw);
    unsigned char* after = smem_raw + (size_t)kBoxBytes * STAGES;
    float* pose_par = reinterpret_cast<float*>(after);                 // [kTmaRound][PP]
    float* pose_acc = pose_par + kTmaRound * PP;                       // [kTmaRound][NV]
    uint64_t* bars = reinterpret_cast<uint64_t*>(pose_acc + kTmaRound * NV + ((kTmaRound * NV) & 1));   // full[S], empty[S]
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    int2* win_org = reinterpret_cast<int2*>(bars + 2 * STAGES);        // [STAGES] window origin (x, y) of the staged pose

    const int pc = blockIdx.x % point_chunks;
    const int64_t bc = blockIdx.x / point_chunks;
    const int64_t b0 = bc * pose_chunk;
    const int n_pose = (int)((b0 + pose_chunk < B ? b0 + pose_chunk : B) - b0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool is_producer = warp == kTmaConsumers / 32;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kTmaConsumers / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (is_producer) {
        // ===== producer warp: window placement + TMA ring ==========================================================
        float c[N_IN];
#pragma unroll
        for (int j = 0; j < N_IN; ++j) c[j] = __ldg(centroid + (int64_t)pc * N_IN + j);
        auto issue = [&](int i) {      // lane 0
            const int s = i % STAGES;
            const int64_t b = b0 + i;
            float proj = 0.f;
#pragma unroll
            for (int j = 0; j < N_IN; ++j) proj = fmaf(__ldg(rotation + b * NR + 1 + 2 * j), c[j], proj);
            const float coord = (proj + 1.f + __ldg(translation + b * 2 + 1)) * grid.scale[1];
            // the lower-corner row of a point at `coord` is about coord - 1: centre the band on it, inside the image
            const float o = floorf(coord) - (float)(wy / 2);
            int oy = (fabsf(o) < 1e9f) ? (int)o : 0;                   // NaN / absurd poses: any band will do
            const int hi = grid.g[1] - wy;
            oy = oy < 0 ? 0 : (oy > hi ? hi : oy);
            win_org[s] = make_int2(0, oy);
            mbar_arrive_expect_tx(&full[s], band_bytes);               // release: the origin is visible to the waiters
            tma_load_1d(reinterpret_cast<unsigned char*>(tiles) + kBoxBytes * s, ds_dout + b * grid.cells + (int64_t)oy * grid.g[0],
                        band_bytes, &full[s]);
        };
        if (lane == 0)
            for (int i = 0; i < STAGES - 1 && i < n_pose; ++i) issue(i);
        __syncwarp();
        for (int i = 0; i < n_pose; ++i) {
            const int j = i + STAGES - 1;           // keep STAGES-1 windows in flight ahead of the consumers
            if (j < n_pose) {
                if (j >= STAGES) mbar_wait(&empty[j % STAGES], ((j / STAGES) - 1) & 1);
                if (lane == 0) issue(j);
                __syncwarp();
            }
        }
        return;
    }

    // ===== consumer warps ======================================================================================
    float x[K][N_IN], pw[K], dpt[K][N_IN], dpw[K];
    unsigned valid_mask = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int p = (pc * K + k) * kTmaConsumers + (int)threadIdx.x;
        const bool valid = p < P;
        valid_mask |= valid ? (1u << k) : 0u;
        const int pp = valid ? p : 0;
        load_point(x[k], points, (int64_t)pp);
        pw[k] = HAS_PW ? __ldg(point_weight + pp) : 1.f;
        if (!valid) {     // padding lanes sit on the origin and never load (see dpr_pullback_tma.cuh)
#pragma unroll
            for (int j = 0; j < N_IN; ++j) x[k][j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < N_IN; ++j) dpt[k][j] = 0.f;
        dpw[k] = 0.f;
    }
    const int g[2] = {grid.g[0], grid.g[1]};
    const float scale[2] = {grid.scale[0], grid.scale[1]};
    const int vsel = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);

    for (int r0 = 0; r0 < n_pose; r0 += kTmaRound) {
        const int n_round = (n_pose - r0 < kTmaRound) ? n_pose - r0 : kTmaRound;
        for (int i = threadIdx.x; i < n_round * PP; i += kTmaConsumers) {
            const int bl = i / PP, v = i % PP;
            const int64_t b = b0 + r0 + bl;
            float val = 0.f;
            if (v < NR) val = __ldg(rotation + b * NR + v);
            else if (v < NR + 2) val = -sub_rn(-1.f, __ldg(translation + b * 2 + (v - NR)));
            else if (v == NR + 2) val = out_weight ? __ldg(out_weight + b) : 1.f;
            pose_par[i] = val;
        }
        for (int i = threadIdx.x; i < n_round * NV; i += kTmaConsumers) pose_acc[i] = 0.f;
        named_bar_sync(1, kTmaConsumers);

        for (int bl = 0; bl < n_round; ++bl) {
            const int i = r0 + bl;
            const int s = i % STAGES;
            float par[PP];
#pragma unroll
            for (int q = 0; q < PP / 4; ++q) {
                const float4 v = reinterpret_cast<const float4*>(pose_par + bl * PP)[q];
                par[4 * q] = v.x; par[4 * q + 1] = v.y; par[4 * q + 2] = v.z; par[4 * q + 3] = v.w;
            }
            float R[2][N_IN];
#pragma unroll
            for (int j = 0; j < N_IN; ++j) { R[0][j] = par[2 * j]; R[1][j] = par[2 * j + 1]; }
            const float neg_origin[2] = {par[NR], par[NR + 1]};
            const float ow = par[NR + 2];
            const float ows[2] = {ow * scale[0], ow * scale[1]};
            const float* tile = reinterpret_cast<const float*>(reinterpret_cast<unsigned char*>(tiles) + kBoxBytes * s);
            const float* img = ds_dout + (b0 + i) * grid.cells;
            asm volatile("" : "+l"(img));               // one opaque base register pair for the fallback loads
            mbar_wait(&full[s], (i / STAGES) & 1);
            const int2 org = win_org[s];
            const float* band = tile - (int64_t)org.y * g[0];      // band[iy * g0 + ix] is cell (ix, iy) for rows in the band

            float acc[8], acc_ow = 0.f;
#pragma unroll
            for (int v = 0; v < 8; ++v) acc[v] = 0.f;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                int ix, iy;
                float dl[2];
                stencil2<float, N_IN>(x[k], R, neg_origin, scale, g, ix, iy, dl);
                const bool valid = (valid_mask >> k) & 1u;
                // ONE branch-free load sequence for both sources: the generic base pointer is the staged band when the
                // whole stencil lies inside it and inside the image horizontally (no bounds predicates needed then),
                // else the pose image in global memory with the per-corner bounds rule (src/raster_pullback.jl:51).
                // A divergent "if (in band) LDS else LDG" would serialise the fallback's L2 latency per point.
                const bool inw = (unsigned)(iy - org.y) < (unsigned)(wy - 1) && (unsigned)ix < (unsigned)(g[0] - 1);
                const float* base = (inw ? band : img) + (iy * g[0] + ix);
                const int row = g[0];
                const bool x_lo = valid && (inw || (unsigned)ix < (unsigned)g[0]);
                const bool x_hi = valid && (inw || (unsigned)(ix + 1) < (unsigned)g[0]);
                const bool y_lo = inw || (unsigned)iy < (unsigned)g[1], y_hi = inw || (unsigned)(iy + 1) < (unsigned)g[1];
                float G00 = 0.f, G10 = 0.f, G01 = 0.f, G11 = 0.f;
                if (x_lo && y_lo) G00 = base[0];
                if (x_hi && y_lo) G10 = base[1];
                if (x_lo && y_hi) G01 = base[row];
                if (x_hi && y_hi) G11 = base[row + 1];
                float s_, gx, gy;
                bilinear_with_gradient(G00, G10, G01, G11, dl[0], dl[1], s_, gx, gy);
                acc_ow += HAS_PW ? s_ * pw[k] : s_;
                dpw[k] += s_ * ow;
                const float sx = gx * (HAS_PW ? ows[0] * pw[k] : ows[0]), sy = gy * (HAS_PW ? ows[1] * pw[k] : ows[1]);
                if constexpr (N_IN == 3) { acc[6] += sx; acc[7] += sy; } else { acc[4] += sx; acc[5] += sy; }
#pragma unroll
                for (int j = 0; j < N_IN; ++j) {
                    acc[2 * j] += sx * x[k][j];
                    acc[2 * j + 1] += sy * x[k][j];
                    dpt[k][j] = fmaf(R[1][j], sy, fmaf(R[0][j], sx, dpt[k][j]));
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);          // this warp is done with the staged window
            if constexpr (N_IN == 2) acc[6] = acc_ow;
            butterfly8(acc, lane);
            if constexpr (N_IN == 3) acc_ow = warp_sum(acc_ow);
            {   // one shared-memory atomicAdd instruction (a CAS loop) for all nine sums: see dpr_pullback_fast.cuh
                int slot = vsel;
                float val = acc[0];
                bool mine = (lane & 3) == 0 && vsel < NV;
                if constexpr (N_IN == 3) {
                    if (lane == 1) { slot = NV - 1; val = acc_ow; mine = true; }
                }
                if (mine) atomicAdd(&pose_acc[bl * NV + slot], val);
            }
        }
        named_bar_sync(1, kTmaConsumers);
        for (int i = threadIdx.x; i < n_round * NV; i += kTmaConsumers) {
            const int bl = i / NV, v = i % NV;
            const float r = pose_acc[i];
            const int64_t b = b0 + r0 + bl;
            if (v < NR) red_add(d_rotation + b * NR + v, r);
            else if (v < NR + 2) red_add(d_translation + b * 2 + (v - NR), r);
            else if (d_out_weight) red_add(d_out_weight + b, r);
        }
        named_bar_sync(1, kTmaConsumers);      // pose_par / pose_acc are rewritten by the next round
    }

#pragma unroll
    for (int k = 0; k < K; ++k) {
        int p = (pc * K + k) * kTmaConsumers + (int)threadIdx.x;
        if (p >= P) continue;
        if (perm) p = __ldg(perm + p);
#pragma unroll
        for (int j = 0; j < N_IN; ++j) red_add(d_points + (int64_t)p * N_IN + j, dpt[k][j]);
        if (d_point_weight) red_add(d_point_weight + p, dpw[k]);
    }
}

inline size_t win_pullback_smem(int64_t g0, int wy, int stages, int n_in) {
    const int NV = 2 * n_in + 3, PP = (NV + 3) / 4 * 4;
    const size_t stage = ((size_t)wy * (size_t)g0 * 4 + 127) / 128 * 128;
    return stage * stages + sizeof(float) * kTmaRound * (PP + NV + 1) + 24 * stages + 64;
}

}  // namespace dpr
