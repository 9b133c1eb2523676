// dpr_pullback_tma.cuh - 2-d Float32 pullback for pose images that fit in shared memory (included by dpr_pullback.cu).
//
// When a whole ds_dout pose image is at most ~64 KB (e.g. 128 x 128 Float32, BASELINE config 5) it is cheaper to bring
// it on chip once per (CTA, pose) than to gather it through L1: a dedicated producer warp streams the images of the
// CTA's pose chunk into a 3-stage shared-memory ring with the TMA unit (cp.async.bulk, one instruction per 64 KB
// image, completion counted on an mbarrier), 15 consumer warps gather with LDS (2.6 T gathers/s vs 0.8 T/s through
// L1, profiles/probe_atomics_r01.json) and release the stage through an "empty" mbarrier.  A thread still OWNS its
// K points across all poses, so d_points needs one REDG per point and pose chunk.  While the consumers work, the
// otherwise idle producer warp sums the staged image: d_background comes from the SAME read of ds_dout
// (ds_dout crosses HBM once), replacing the separate background_sum pass.
#pragma once
#include <cuda.h>
#include "dpr_common.cuh"
#include "dpr_pullback_fast.cuh"  // stencil2, butterfly8

namespace dpr {

constexpr int kTmaConsumers = 480;      // consumer threads (15 warps) + 32 producer threads = 512 (128 registers each)
constexpr int kTmaRound = 64;           // poses whose parameters / accumulators are resident at a time

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// PADDED: the image is fetched with ONE 3-d tensor-map TMA copy (cp.async.bulk.tensor, SASS UTMALDG) whose box is four
// columns WIDER than the image: the extra columns lie outside the tensor, arrive as zeros, and give the staged image a row
// pitch of g0 + 4 words.  With the dense pitch (a multiple of the 32 banks) the bank of a cell depends on its column
// only, and spatially sorted lanes - a blob ~10 pixels wide under every pose - serialised 5.3 ways per LDS (66 % of the
// kernel's shared-memory wavefronts were bank conflicts, profiles/ncu_full_r01_v10_cfg5_summary.csv); with the padded pitch
// the bank is (x + 4 y) mod 32.
// The box also starts four columns LEFT of the image and two rows ABOVE it and ends two rows below: the staged image sits in
// a frame of zeros - (row r, column c) is word (r + 2) * pitch + c + 4, columns g0 and g0 + 1 of a row are the zero padding in
// front of the next row (four guard words, zeroed once, follow the last row) - so the consumers clamp the lower-corner cell
// into [-2, g] and load all four corners WITHOUT bounds predicates (src/raster_pullback.jl:51 skips out-of-bounds corners;
// here they read zeros): 13 instructions per splat instead of 18.
// !PADDED: 1-d bulk copy of the dense image (cp.async.bulk, UBLKCP), for rows that are not multiples of 16 bytes.
template <int N_IN, int K, bool HAS_PW, int STAGES, bool PADDED>
__global__ void __launch_bounds__(kTmaConsumers + 32, 1)
pullback_tma2d_kernel(const __grid_constant__ CUtensorMap map, const float* __restrict__ ds_dout, const float* __restrict__ points,
                      const float* __restrict__ rotation, const float* __restrict__ translation,
                      const float* __restrict__ out_weight, const float* __restrict__ point_weight,
                      float* __restrict__ d_points, float* __restrict__ d_rotation, float* __restrict__ d_translation,
                      float* __restrict__ d_background, float* __restrict__ d_out_weight,
                      float* __restrict__ d_point_weight, const int32_t* __restrict__ perm, Grid<float, 2> grid, int P,
                      int64_t B, int point_chunks, int pose_chunk) {
    constexpr int NR = 2 * N_IN, NV = NR + 3, PP = (NV + 3) / 4 * 4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int pitch = PADDED ? grid.g[0] + 4 : grid.g[0];            // words per staged row
    const int cells = (int)grid.cells;
    const int stage_words = pitch * (PADDED ? grid.g[1] + 4 : grid.g[1]);     // PADDED: rows -2 .. g1 + 1
    const uint32_t img_bytes = (uint32_t)stage_words * 4u;           // bytes one copy delivers (zero-filled frame included)
    const size_t stage_stride = ((size_t)img_bytes + (PADDED ? 16 : 0) + 127) / 128 * 128;      // as tma_pullback_smem()
    float* tiles = reinterpret_cast<float*>(smem_raw);
    unsigned char* after = smem_raw + stage_stride * STAGES;
    float* pose_par = reinterpret_cast<float*>(after);                 // [kTmaRound][PP]
    float* pose_acc = pose_par + kTmaRound * PP;                       // [kTmaRound][NV]
    uint64_t* bars = reinterpret_cast<uint64_t*>(pose_acc + kTmaRound * NV + ((kTmaRound * NV) & 1));   // full[S], empty[S]
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;

    const int pc = blockIdx.x % point_chunks;
    const int64_t bc = blockIdx.x / point_chunks;
    const int64_t b0 = bc * pose_chunk;
    const int n_pose = (int)((b0 + pose_chunk < B ? b0 + pose_chunk : B) - b0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool is_producer = warp == kTmaConsumers / 32;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kTmaConsumers / 32 + 1);
        }
        mbar_fence_init();
        if constexpr (PADDED) {          // guard words behind each stage's last row (read as columns g0, g0 + 1 of row g1 + 1)
            for (int s = 0; s < STAGES; ++s)
                for (int w = 0; w < 4; ++w) reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(tiles) + stage_stride * s)[stage_words + w] = 0.f;
        }
    }
    __syncthreads();

    if (is_producer) {
        // ===== producer warp: TMA ring + d_background ==========================================================
        const bool do_bg = d_background != nullptr && pc == 0;
        auto issue = [&](int i) {      // lane 0 only
            const int s = i % STAGES;
            mbar_arrive_expect_tx(&full[s], img_bytes);
            if constexpr (PADDED) {
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                                 smem_u32(reinterpret_cast<unsigned char*>(tiles) + stage_stride * s)),
                             "l"(&map), "r"(-4), "r"(-2), "r"((int)(b0 + i)), "r"(smem_u32(&full[s]))
                             : "memory");
            } else {
                tma_load_1d(reinterpret_cast<unsigned char*>(tiles) + stage_stride * s, ds_dout + (b0 + i) * (int64_t)cells,
                            img_bytes, &full[s]);
            }
        };
        if (lane == 0)
            for (int i = 0; i < STAGES - 1 && i < n_pose; ++i) issue(i);
        for (int i = 0; i < n_pose; ++i) {
            const int j = i + STAGES - 1;           // keep STAGES-1 images in flight ahead of the consumers
            if (j < n_pose) {
                if (j >= STAGES) mbar_wait(&empty[j % STAGES], ((j / STAGES) - 1) & 1);
                if (lane == 0) issue(j);
            }
            const int s = i % STAGES;
            if (do_bg) {
                mbar_wait(&full[s], (i / STAGES) & 1);
                const float4* t4 = reinterpret_cast<const float4*>(reinterpret_cast<unsigned char*>(tiles) + stage_stride * s);
                float acc0 = 0.f, acc1 = 0.f;
                for (int q = lane; q < stage_words / 4; q += 32) {     // (the padding columns are zeros)
                    const float4 v = t4[q];
                    acc0 += v.x + v.y;
                    acc1 += v.z + v.w;
                }
                const float tot = warp_sum(acc0 + acc1);
                if (lane == 0) d_background[b0 + i] = tot;      // src/raster_pullback.jl:78
            }
            __syncwarp();           // (the store of the tile sum above depends on every lane's loads: they have returned)
            if (lane == 0) mbar_arrive(&empty[s]);
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
        if (!valid) {
            // padding lanes sit on the origin and never load (corner predicates include the valid bit): exact zeros.
            // A far-away point is not safe - a matrix row orthogonal to it projects it into the image.
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
        // stage this round's pose parameters, clear its accumulators (consumer warps only: named barrier 1)
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
            const float* __restrict__ tile = reinterpret_cast<const float*>(reinterpret_cast<unsigned char*>(tiles) + stage_stride * s);
            mbar_wait(&full[s], (i / STAGES) & 1);
            const float* __restrict__ frame = tile + (2 * pitch + 4);   // PADDED: cell (0, 0) of the framed image

            float acc[8], acc_ow = 0.f;
#pragma unroll
            for (int v = 0; v < 8; ++v) acc[v] = 0.f;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                int ix, iy;
                float dl[2];
                stencil2<float, N_IN>(x[k], R, neg_origin, scale, g, ix, iy, dl);
                const bool valid = (valid_mask >> k) & 1u;
                float G00 = 0.f, G10 = 0.f, G01 = 0.f, G11 = 0.f;
                if constexpr (PADDED) {
                    // clamp into the frame of zeros (at -2 both corners of a dimension are outside the image); padding lanes
                    // (p >= P) are sent to the zero columns
                    int cx = ix < -2 ? -2 : ix, cy = iy < -2 ? -2 : iy;
                    cx = cx > g[0] ? g[0] : cx;
                    cy = cy > g[1] ? g[1] : cy;
                    if (!valid) cx = g[0];
                    const float* __restrict__ c00 = frame + (cy * pitch + cx);
                    G00 = c00[0];
                    G10 = c00[1];
                    G01 = c00[pitch];
                    G11 = c00[pitch + 1];
                } else {
                    const bool x_lo = valid && (unsigned)ix < (unsigned)g[0], x_hi = valid && (unsigned)(ix + 1) < (unsigned)g[0];
                    const bool y_lo = (unsigned)iy < (unsigned)g[1], y_hi = (unsigned)(iy + 1) < (unsigned)g[1];
                    const int off = iy * pitch + ix;
                    if (x_lo && y_lo) G00 = tile[off];
                    if (x_hi && y_lo) G10 = tile[off + 1];
                    if (x_lo && y_hi) G01 = tile[off + pitch];
                    if (x_hi && y_hi) G11 = tile[off + pitch + 1];
                }
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
            // Release the stage HERE, not right after the gathers.  A warp issues in order and an instruction waits for its
            // operands: the atomic above cannot issue before `val` - which depends, through the shuffles, on every load of
            // every lane from the stage - is ready, so by now all of this warp's reads of the staged image have returned.
            // An arrive straight after the point loop issues while the last loads are still in flight (SASS: LDS x 4,
            // WARPSYNC, SYNCS.ARRIVE; their consumers are scheduled behind it), and the producer's next TMA copy into the
            // stage then overwrote cells a late lane was about to read: a handful of wrong splats of one pose in one warp, in
            // one run out of three on some shapes (tools/soak.py; tests/test_gpu_parity.py::
            // test_pullback_tma2d_stage_release_is_ordered).  Measured alternatives on config 5 (10 points per thread):
            // fence.acq_rel.cta in front of the arrive 5.92 ms, a volatile store of acc_ow in front of it 5.83 ms, this 5.66 ms
            // (the racy release: 5.43 ms); with 8 points per thread this release runs at 5.55 ms.
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
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

// `cells`: words one stage holds - g0 * g1 (dense) or (g0 + 4) * (g1 + 4) + 4 guard words (padded)
inline size_t tma_pullback_smem(int64_t cells, int stages, int n_in) {
    const int NV = 2 * n_in + 3, PP = (NV + 3) / 4 * 4;
    const size_t stage_stride = ((size_t)cells * 4 + 127) / 128 * 128;
    return stage_stride * stages + sizeof(float) * kTmaRound * (PP + NV + 1) + 16 * stages + 64;
}

}  // namespace dpr
