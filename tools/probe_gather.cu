// Microbenchmark: which load path is fastest for the pullback's 2x2 gathers from an L2-resident 256x256 image?
// __ldg (ld.global.nc), ld.global.ca, ld.global.cg (bypass L1), tex1Dfetch through a linear texture object.
// Lanes of a warp land in a blob of `blob` x `blob` pixels (blob = 255: random), like the spatially sorted points.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/probe_gather tools/probe_gather.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__device__ __forceinline__ float ld_ca(const float* p) { float v; asm volatile("ld.global.ca.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; }
__device__ __forceinline__ float ld_cg(const float* p) { float v; asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; }

template <int MODE>
__global__ void __launch_bounds__(256) k_gather(const float* __restrict__ img, cudaTextureObject_t tex, float* sink, int G, int n_img, int iters, int blob) {
    float acc = 0.f;
    uint32_t seed = blockIdx.x * 256u + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        const int pose = (blockIdx.x / 8 + it) % n_img;          // neighbouring CTAs work on the same image at the same time
        const uint32_t h = hash32(seed + it * 0x9e3779b9U), hw = hash32((seed >> 5) + it * 0x9e3779b9U);
        const int x = (hw % (G - blob)) + (h % blob), y = ((hw >> 12) % (G - blob)) + ((h >> 10) % blob);
        const int64_t base = (int64_t)pose * G * G + (int64_t)y * G + x;
        const float w = (float)(h >> 24) * (1.f / 256.f);
        float g00, g10, g01, g11;
        if (MODE == 0) { g00 = __ldg(img + base); g10 = __ldg(img + base + 1); g01 = __ldg(img + base + G); g11 = __ldg(img + base + G + 1); }
        else if (MODE == 1) { g00 = ld_ca(img + base); g10 = ld_ca(img + base + 1); g01 = ld_ca(img + base + G); g11 = ld_ca(img + base + G + 1); }
        else if (MODE == 2) { g00 = ld_cg(img + base); g10 = ld_cg(img + base + 1); g01 = ld_cg(img + base + G); g11 = ld_cg(img + base + G + 1); }
        else if (MODE == 3) { g00 = tex1Dfetch<float>(tex, (int)base); g10 = tex1Dfetch<float>(tex, (int)base + 1); g01 = tex1Dfetch<float>(tex, (int)base + G); g11 = tex1Dfetch<float>(tex, (int)base + G + 1); }
        else if (MODE == 4) { g00 = __ldg(img + base); g10 = __ldg(img + base + 1); g01 = __ldg(img + base + G); g11 = tex1Dfetch<float>(tex, (int)base + G + 1); }   // 3 LSU + 1 TEX
        else { g00 = __ldg(img + base); g10 = __ldg(img + base + 1); g01 = tex1Dfetch<float>(tex, (int)base + G); g11 = tex1Dfetch<float>(tex, (int)base + G + 1); }   // 2 LSU + 2 TEX
        acc += g00 * w + g10 * (1.f - w) + g01 * w + g11;
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
    const int G = 256, n_img = 256, iters = 1024;     // 64 MB of images: L2 resident
    float* img; CK(cudaMalloc(&img, (size_t)n_img * G * G * 4)); CK(cudaMemset(img, 0, (size_t)n_img * G * G * 4));
    float* sink; CK(cudaMalloc(&sink, 1024));
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = img;
    rd.res.linear.desc = cudaCreateChannelDesc<float>(); rd.res.linear.sizeInBytes = (size_t)n_img * G * G * 4;
    cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex = 0; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    printf("{\"device\": \"%s\", \"results\": [\n", p.name);
    const int ctas = p.multiProcessorCount * 8;
    const char* names[6] = {"ldg_nc", "ld_ca", "ld_cg", "tex1Dfetch", "3ldg_1tex", "2ldg_2tex"};
    for (int blob : {255, 26, 8}) {
        for (int mode = 0; mode < 6; ++mode) {
            double ms = 0;
            if (mode == 0) ms = time_ms([&] { k_gather<0><<<ctas, 256>>>(img, tex, sink, G, n_img, iters, blob); }, 3);
            if (mode == 1) ms = time_ms([&] { k_gather<1><<<ctas, 256>>>(img, tex, sink, G, n_img, iters, blob); }, 3);
            if (mode == 2) ms = time_ms([&] { k_gather<2><<<ctas, 256>>>(img, tex, sink, G, n_img, iters, blob); }, 3);
            if (mode == 3) ms = time_ms([&] { k_gather<3><<<ctas, 256>>>(img, tex, sink, G, n_img, iters, blob); }, 3);
            if (mode == 4) ms = time_ms([&] { k_gather<4><<<ctas, 256>>>(img, tex, sink, G, n_img, iters, blob); }, 3);
            if (mode == 5) ms = time_ms([&] { k_gather<5><<<ctas, 256>>>(img, tex, sink, G, n_img, iters, blob); }, 3);
            printf(" {\"path\": \"%s\", \"blob\": %d, \"ms\": %.4f, \"corner_loads_per_s\": %.4e},\n", names[mode], blob, ms, (double)ctas * 256 * iters * 4 / (ms * 1e-3));
            fflush(stdout);
        }
    }
    printf(" {}]}\n");
    return 0;
}
