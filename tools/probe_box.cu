// Microbenchmark for a pullback variant that was estimated but not built: per (CTA, pose) stage the 64 x 64
// pixel box that holds the projected blob of the CTA's 1024 spatially sorted points in shared memory (cp.async, two
// stages, skewed pitch) and gather the 2 x 2 stencils with LDS, against gathering them directly from L1 / L2 with __ldg as
// pullback_gather2d_kernel does.  Geometry like config 2: 256 x 256 images, L2 resident; a CTA's points fall into a
// 56-pixel blob, the 32 lanes of a warp into a `warp_blob`-pixel blob inside it (26 for 100 k points in 3-d).
// Reports corner loads per second (box loads included).  Occupancy is set through the dynamic shared memory size
// (3 CTAs per SM like the real kernel with its 77 registers, and 6).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/probe_box tools/probe_box.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

constexpr int G = 256, BOX = 64, CTA_BLOB = 56, K = 4;

// CTA blob origin (ox, oy) for "pose" it: the same for the whole CTA
__device__ __forceinline__ void cta_origin(int it, int& ox, int& oy) {
    const uint32_t h = hash32(blockIdx.x * 0x85ebca6bU + it * 0x9e3779b9U);
    ox = h % 190; oy = (h >> 12) % 190;
}
// Relative position of point k of this thread inside the CTA blob, fixed for the whole run like the points a thread owns
// (computed once, outside the timed loop's critical path; per "pose" only the CTA origin moves, so the loop has no
// division or hashing per point and both variants are bound by their memory path, not by integer arithmetic).
__device__ __forceinline__ void point_rel(int k, int warp_blob, int& rx, int& ry, float& w) {
    const uint32_t warp_id = (blockIdx.x * 8u + (threadIdx.x >> 5)) * K + k;
    const uint32_t hw = hash32(warp_id * 0xc2b2ae35U + 0x9e3779b9U);
    const uint32_t hl = hash32((blockIdx.x * 256u + threadIdx.x) * K + k + 0x27d4eb2fU);
    rx = (int)(hw % (CTA_BLOB - warp_blob)) + (int)(hl % warp_blob);
    ry = (int)((hw >> 12) % (CTA_BLOB - warp_blob)) + (int)((hl >> 10) % warp_blob);
    w = (float)(hl >> 24) * (1.f / 256.f);
}

__global__ void __launch_bounds__(256) k_direct(const float* __restrict__ img, float* sink, int n_img, int iters, int warp_blob) {
    float acc = 0.f;
    int rx[K], ry[K]; float wt[K];
#pragma unroll
    for (int k = 0; k < K; ++k) point_rel(k, warp_blob, rx[k], ry[k], wt[k]);
    for (int it = 0; it < iters; ++it) {
        const int pose = (blockIdx.x / 8 + it) % n_img;          // neighbouring CTAs work on the same image at the same time
        int ox, oy;
        cta_origin(it, ox, oy);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int x = ox + rx[k], y = oy + ry[k]; const float w = wt[k];
            const float* p = img + (int64_t)pose * G * G + (int64_t)y * G + x;
            const float g00 = __ldg(p), g10 = __ldg(p + 1), g01 = __ldg(p + G), g11 = __ldg(p + G + 1);
            acc += g00 * w + g10 * (1.f - w) + g01 * w + g11;
        }
    }
    if (acc == 123.456f) sink[0] = acc;
}

template <int PITCH>
__global__ void __launch_bounds__(256) k_box(const float* __restrict__ img, float* sink, int n_img, int iters, int warp_blob) {
    extern __shared__ __align__(16) float sm[];                   // [2][BOX][PITCH]
    float acc = 0.f;
    int rx[K], ry[K]; float wt[K];
#pragma unroll
    for (int k = 0; k < K; ++k) point_rel(k, warp_blob, rx[k], ry[k], wt[k]);
    auto issue = [&](int it) {
        const int pose = (blockIdx.x / 8 + it) % n_img;
        int ox, oy;
        cta_origin(it, ox, oy);
        const int bx = ox & ~3;                                   // 16-byte aligned box origin
        const float* src0 = img + (int64_t)pose * G * G + (int64_t)oy * G + bx;
        float* dst0 = sm + (it & 1) * BOX * PITCH;
#pragma unroll
        for (int c = threadIdx.x; c < BOX * (BOX / 4); c += 256) {
            const int row = c / (BOX / 4), col = (c % (BOX / 4)) * 4;
            const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst0 + row * PITCH + col);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src0 + (int64_t)row * G + col) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    issue(0);
    for (int it = 0; it < iters; ++it) {
        if (it + 1 < iters) {
            issue(it + 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        int ox, oy;
        cta_origin(it, ox, oy);
        const int bx = ox & ~3;
        const float* tile = sm + (it & 1) * BOX * PITCH;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int x = ox + rx[k], y = oy + ry[k]; const float w = wt[k];
            const float* p = tile + (y - oy) * PITCH + (x - bx);
            const float g00 = p[0], g10 = p[1], g01 = p[PITCH], g11 = p[PITCH + 1];
            acc += g00 * w + g10 * (1.f - w) + g01 * w + g11;
        }
        __syncthreads();                                          // the stage is overwritten by the copy issued next iteration
    }
    if (acc == 123.456f) sink[0] = acc;
}

template <typename F>
static double time_ms(F launch, int reps) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    launch(); launch(); CK(cudaDeviceSynchronize()); CK(cudaGetLastError());
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) launch();
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); CK(cudaGetLastError());
    return ms / reps;
}

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int n_img = 256, iters = 1024;                          // 64 MB of images: L2 resident
    float* img; CK(cudaMalloc(&img, (size_t)n_img * G * G * 4)); CK(cudaMemset(img, 0, (size_t)n_img * G * G * 4));
    float* sink; CK(cudaMalloc(&sink, 1024));
    CK(cudaFuncSetAttribute(k_direct, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CK(cudaFuncSetAttribute(k_box<68>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CK(cudaFuncSetAttribute(k_box<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    printf("{\"device\": \"%s\", \"box\": %d, \"cta_blob\": %d, \"points_per_thread\": %d, \"results\": [\n", p.name, BOX, CTA_BLOB, K);
    const int ctas = p.multiProcessorCount * 12;
    for (int occ : {3, 6}) {
        const size_t smem = occ == 3 ? 74 * 1024 : 36 * 1024;     // 227 KB / occ, >= the 34.8 KB the box kernel needs
        for (int blob : {26, 12}) {
            double ms;
            const double loads = (double)ctas * 256 * iters * K * 4;
            ms = time_ms([&] { k_direct<<<ctas, 256, smem>>>(img, sink, n_img, iters, blob); }, 3);
            printf(" {\"path\": \"direct_ldg\", \"ctas_per_sm\": %d, \"warp_blob\": %d, \"ms\": %.4f, \"corner_loads_per_s\": %.4e},\n", occ, blob, ms, loads / (ms * 1e-3));
            ms = time_ms([&] { k_box<68><<<ctas, 256, smem>>>(img, sink, n_img, iters, blob); }, 3);
            printf(" {\"path\": \"box_cp_async_pitch68\", \"ctas_per_sm\": %d, \"warp_blob\": %d, \"ms\": %.4f, \"corner_loads_per_s\": %.4e},\n", occ, blob, ms, loads / (ms * 1e-3));
            ms = time_ms([&] { k_box<64><<<ctas, 256, smem>>>(img, sink, n_img, iters, blob); }, 3);
            printf(" {\"path\": \"box_cp_async_pitch64\", \"ctas_per_sm\": %d, \"warp_blob\": %d, \"ms\": %.4f, \"corner_loads_per_s\": %.4e},\n", occ, blob, ms, loads / (ms * 1e-3));
            fflush(stdout);
        }
    }
    printf(" {}]}\n");
    return 0;
}
