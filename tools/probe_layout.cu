// Microbenchmark: would a TILED layout of ds_dout make the pullback's 2x2 gathers cheaper on the L1 data pipe?
// Same access pattern as tools/probe_gather.cu (the lanes of a warp land in a blob x blob pixel neighbourhood of an
// L2-resident 256 x 256 image), four layouts / load shapes:
//   row_scalar   row-major image, 4 x LDG.32
//   row_pair     row-major image, aligned LDG.64 per row + LDG.32 for odd columns (what pullback_gather2d_kernel does)
//   tile8x4      8 x 4-pixel tiles (one 128-byte line each), 4 x LDG.32
//   tile4x4p     4 x 4-pixel tiles (64 bytes), rows of a tile contiguous: aligned LDG.64 pairs inside the tile where possible
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/probe_layout tools/probe_layout.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <int TW, int TH>
__device__ __forceinline__ int tiled(int x, int y, int G) { return ((y / TH) * (G / TW) + (x / TW)) * (TW * TH) + (y % TH) * TW + (x % TW); }

template <int MODE>
__global__ void __launch_bounds__(256) k_gather(const float* __restrict__ img, float* sink, int G, int n_img, int iters, int blob) {
    float acc = 0.f;
    uint32_t seed = blockIdx.x * 256u + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        const int pose = (blockIdx.x / 8 + it) % n_img;
        const uint32_t h = hash32(seed + it * 0x9e3779b9U), hw = hash32((seed >> 5) + it * 0x9e3779b9U);
        const int x = (hw % (G - blob - 1)) + (h % blob), y = ((hw >> 12) % (G - blob - 1)) + ((h >> 10) % blob);
        const float* im = img + (int64_t)pose * G * G;
        const float w = (float)(h >> 24) * (1.f / 256.f);
        float g00, g10, g01, g11;
        if (MODE == 0) {
            const int b = y * G + x;
            g00 = __ldg(im + b); g10 = __ldg(im + b + 1); g01 = __ldg(im + b + G); g11 = __ldg(im + b + G + 1);
        } else if (MODE == 1) {
            const int b = y * G + x; const bool odd = x & 1;
            const float2 q0 = __ldg(reinterpret_cast<const float2*>(im + (b & ~1))), q1 = __ldg(reinterpret_cast<const float2*>(im + (b & ~1) + G));
            float e0 = 0.f, e1 = 0.f;
            if (odd) { e0 = __ldg(im + (b & ~1) + 2); e1 = __ldg(im + (b & ~1) + G + 2); }
            g00 = odd ? q0.y : q0.x; g10 = odd ? e0 : q0.y; g01 = odd ? q1.y : q1.x; g11 = odd ? e1 : q1.y;
        } else if (MODE == 2) {
            g00 = __ldg(im + tiled<8, 4>(x, y, G)); g10 = __ldg(im + tiled<8, 4>(x + 1, y, G));
            g01 = __ldg(im + tiled<8, 4>(x, y + 1, G)); g11 = __ldg(im + tiled<8, 4>(x + 1, y + 1, G));
        } else {
            // 4x4 tiles: the x-pair is one aligned LDG.64 when x is even (same tile row); odd x: two scalar loads
            const bool odd = x & 1;
            float2 q0 = make_float2(0.f, 0.f), q1 = make_float2(0.f, 0.f);
            if (!odd) { q0 = __ldg(reinterpret_cast<const float2*>(im + tiled<4, 4>(x, y, G))); q1 = __ldg(reinterpret_cast<const float2*>(im + tiled<4, 4>(x, y + 1, G))); }
            else { q0.x = __ldg(im + tiled<4, 4>(x, y, G)); q0.y = __ldg(im + tiled<4, 4>(x + 1, y, G));
                   q1.x = __ldg(im + tiled<4, 4>(x, y + 1, G)); q1.y = __ldg(im + tiled<4, 4>(x + 1, y + 1, G)); }
            g00 = q0.x; g10 = q0.y; g01 = q1.x; g11 = q1.y;
        }
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
    const int G = 256, n_img = 256, iters = 1024;
    float* img; CK(cudaMalloc(&img, (size_t)n_img * G * G * 4)); CK(cudaMemset(img, 0, (size_t)n_img * G * G * 4));
    float* sink; CK(cudaMalloc(&sink, 1024));
    printf("{\"device\": \"%s\", \"results\": [\n", p.name);
    const int ctas = p.multiProcessorCount * 8;
    const char* names[4] = {"row_scalar", "row_pair", "tile8x4_scalar", "tile4x4_pair"};
    bool first = true;
    for (int blob : {40, 26, 12, 8}) for (int mode = 0; mode < 4; ++mode) {
        double ms = 0;
        if (mode == 0) ms = time_ms([&] { k_gather<0><<<ctas, 256>>>(img, sink, G, n_img, iters, blob); }, 3);
        if (mode == 1) ms = time_ms([&] { k_gather<1><<<ctas, 256>>>(img, sink, G, n_img, iters, blob); }, 3);
        if (mode == 2) ms = time_ms([&] { k_gather<2><<<ctas, 256>>>(img, sink, G, n_img, iters, blob); }, 3);
        if (mode == 3) ms = time_ms([&] { k_gather<3><<<ctas, 256>>>(img, sink, G, n_img, iters, blob); }, 3);
        printf("%s {\"layout\": \"%s\", \"blob\": %d, \"ms\": %.4f, \"corner_loads_per_s\": %.4e}", first ? "" : ",\n", names[mode], blob, ms,
               (double)ctas * 256 * iters * 4 / (ms * 1e-3));
        first = false; fflush(stdout);
    }
    printf("\n]}\n");
    return 0;
}
