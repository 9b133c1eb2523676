// Microbenchmark: how fast can 148 pose-owning CTAs stream COALESCED global reductions into one small (1.6 MB) d_points
// buffer?  This decides whether a pullback with the pose image in shared memory (CTA owns a pose, every thread visits
// ~98 points per pose) can afford one REDG per (point, pose, component) instead of register accumulation over poses.
// Layouts: aos3 (x y z per point, 3 x REDG.F32, lanes 12 B apart), soa4 (4 planes, 4 x REDG.F32, lanes contiguous),
// aos4 (float4 per point, 4 x REDG.F32), aos4_v2 (2 x red.v2.f32), aos4_v4 (1 x red.v4.f32).
// `stagger`: CTAs start their sweep over the points at different offsets (as desynchronised CTAs would).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/probe_redg tools/probe_redg.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void red1(float* p, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void red2(float* p, float a, float b) { asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory"); }
__device__ __forceinline__ void red4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_redg(float* __restrict__ d, int P, int iters, int stagger, int alu) {
    const int n_chunk = (P + 1023) / 1024;
    for (int it = 0; it < iters; ++it) {
        const int start = stagger ? (int)((blockIdx.x * 37u + it * 11u) % n_chunk) : 0;
        for (int c = 0; c < n_chunk; ++c) {
            int cc = c + start; if (cc >= n_chunk) cc -= n_chunk;
            const int p = cc * 1024 + threadIdx.x;
            if (p >= P) continue;
            float v = (float)(p & 255) * 1e-3f + (float)it;
            for (int a = 0; a < alu; ++a) v = fmaf(v, 1.0001f, 0.5f);     // stand-in for the stencil arithmetic
            if (MODE == 0) { red1(d + 3 * (size_t)p, v); red1(d + 3 * (size_t)p + 1, v + 1.f); red1(d + 3 * (size_t)p + 2, v + 2.f); }
            else if (MODE == 1) { red1(d + p, v); red1(d + P + p, v + 1.f); red1(d + 2 * (size_t)P + p, v + 2.f); red1(d + 3 * (size_t)P + p, v + 3.f); }
            else if (MODE == 2) { float* q = d + 4 * (size_t)p; red1(q, v); red1(q + 1, v + 1.f); red1(q + 2, v + 2.f); red1(q + 3, v + 3.f); }
            else if (MODE == 3) { float* q = d + 4 * (size_t)p; red2(q, v, v + 1.f); red2(q + 2, v + 2.f, v + 3.f); }
            else if (MODE == 4) { red4(d + 4 * (size_t)p, v, v + 1.f, v + 2.f, v + 3.f); }
            else if (MODE == 5) { red1(d + p, v); red1(d + P + p, v + 1.f); red1(d + 2 * (size_t)P + p, v + 2.f); }   // soa3
        }
    }
}

template <typename F>
static double time_ms(F launch, int reps) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    launch(); CK(cudaDeviceSynchronize()); CK(cudaGetLastError());
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) launch();
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); CK(cudaGetLastError());
    return ms / reps;
}

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int P = 100000, iters = 28;                 // 148 CTAs x 28 poses x 100k points = config 2's splat count
    float* d; CK(cudaMalloc(&d, (size_t)P * 4 * 4)); CK(cudaMemset(d, 0, (size_t)P * 16));
    const int ctas = prop.multiProcessorCount;
    const char* names[6] = {"aos3_3xf32", "soa4_4xf32", "aos4_4xf32", "aos4_2xv2", "aos4_1xv4", "soa3_3xf32"};
    const int comps[6] = {3, 4, 4, 4, 4, 3};
    printf("{\"device\": \"%s\", \"ctas\": %d, \"points\": %d, \"poses_per_cta\": %d, \"results\": [\n", prop.name, ctas, P, iters);
    bool first = true;
    for (int alu : {0, 60}) for (int stagger : {0, 1}) for (int mode = 0; mode < 6; ++mode) {
        double ms = 0;
        auto go = [&](auto kern) { ms = time_ms([&] { kern<<<ctas, 1024>>>(d, P, iters, stagger, alu); }, 3); };
        switch (mode) {
            case 0: go(k_redg<0>); break; case 1: go(k_redg<1>); break; case 2: go(k_redg<2>); break;
            case 3: go(k_redg<3>); break; case 4: go(k_redg<4>); break; default: go(k_redg<5>); break;
        }
        const double splats = (double)ctas * iters * P;
        printf("%s {\"case\": \"%s\", \"stagger\": %d, \"alu_per_splat\": %d, \"ms\": %.4f, \"splats_per_s\": %.4e, \"lane_ops_per_s\": %.4e}",
               first ? "" : ",\n", names[mode], stagger, alu, ms, splats / (ms * 1e-3), splats * comps[mode] / (ms * 1e-3));
        first = false;
    }
    printf("\n]}\n");
    return 0;
}
