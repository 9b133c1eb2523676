// Microbenchmark 2 for the forward splat: native shared-memory ATOMS.ADD rate of the 4-corner stencil as a function of
//   pitch   row stride of the tile in words (256 = config 2's image: the bank depends on the column only),
//   dist    0 uniform positions, 1 Gaussian blob (sigma = 51 px, config 2's projected cloud), 2 Gaussian with the lanes
//           of a warp on a thin ring (radius-sorted points),
//   swz     0 none, 1 column index XOR-swizzled with the row (bank = f(x, y) even for pitch 256).
// 1 CTA of 1024 threads per SM, 220 rows on chip.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/probe_atoms2 tools/probe_atoms2.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__device__ __forceinline__ float u01(uint32_t h) { return (float)(h >> 8) * (1.f / 16777216.f) + 1e-7f; }

template <int SWZ>
__global__ void __launch_bounds__(1024, 1) k_atoms(float* sink, int cols, int rows, int pitch, int iters, int dist) {
    extern __shared__ unsigned tile[];
    const int n = rows * pitch;
    for (int i = threadIdx.x; i < n; i += blockDim.x) tile[i] = 0u;
    __syncthreads();
    uint32_t seed = blockIdx.x * 1024u + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        const uint32_t h = hash32(seed + it * 0x9e3779b9U), h2 = hash32(h ^ 0x68bc21ebU);
        int x, y;
        if (dist == 0) { x = h % (cols - 1); y = (h >> 12) % (rows - 1); }
        else {
            float r = sqrtf(-2.f * __logf(u01(h)));                   // Rayleigh radius of a 2-d Gaussian
            if (dist == 2) r = 0.02f * r + 2.5f * u01(hash32((seed >> 5) + it * 0x9e3779b9U));   // thin ring per warp
            float sn, cs; __sincosf(6.2831853f * u01(h2), &sn, &cs);
            x = (int)(cols * 0.5f + 51.f * r * cs); y = (int)(rows * 0.5f + 51.f * r * sn);
            x = min(max(x, 0), cols - 2); y = min(max(y, 0), rows - 2);
        }
        const unsigned q = h >> 20;
        if (SWZ == 0) {
            const int base = y * pitch + x;
            atomicAdd(&tile[base], q); atomicAdd(&tile[base + 1], q + 1);
            atomicAdd(&tile[base + pitch], q + 2); atomicAdd(&tile[base + pitch + 1], q + 3);
        } else {
            // swizzle: column ^ ((row & 7) << 2) keeps 4-word groups intact, spreads rows over banks
            const int s0 = (y & 7) << 2, s1 = ((y + 1) & 7) << 2;
            atomicAdd(&tile[y * pitch + (x ^ s0)], q); atomicAdd(&tile[y * pitch + ((x + 1) ^ s0)], q + 1);
            atomicAdd(&tile[(y + 1) * pitch + (x ^ s1)], q + 2); atomicAdd(&tile[(y + 1) * pitch + ((x + 1) ^ s1)], q + 3);
        }
    }
    __syncthreads();
    unsigned s = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += tile[i];
    if (s == 123456789u) sink[0] = (float)s;
}

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    float* sink; CK(cudaMalloc(&sink, 1024));
    const int iters = 2048, ctas = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"results\": [\n", p.name);
    bool first = true;
    struct Case { int cols, rows, pitch, dist, swz; };
    const Case cases[] = {{220, 220, 220, 0, 0}, {256, 220, 256, 0, 0}, {256, 220, 256, 1, 0}, {256, 220, 256, 2, 0},
                          {256, 213, 264, 1, 0}, {256, 213, 264, 2, 0}, {256, 220, 256, 0, 1}, {256, 220, 256, 1, 1},
                          {256, 220, 256, 2, 1}, {256, 216, 260, 1, 0}, {256, 216, 260, 2, 0}};
    for (const Case& c : cases) {
        const size_t smem = (size_t)c.rows * c.pitch * 4;
        auto kern = c.swz ? k_atoms<1> : k_atoms<0>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
        kern<<<ctas, 1024, smem>>>(sink, c.cols, c.rows, c.pitch, iters, c.dist);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(a));
        for (int r = 0; r < 3; ++r) kern<<<ctas, 1024, smem>>>(sink, c.cols, c.rows, c.pitch, iters, c.dist);
        CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b)); ms /= 3; CK(cudaGetLastError());
        const double ops = (double)ctas * 1024.0 * iters * 4.0;
        printf("%s {\"cols\": %d, \"rows\": %d, \"pitch\": %d, \"dist\": %d, \"swizzle\": %d, \"ms\": %.4f, \"corner_ops_per_s\": %.4e}",
               first ? "" : ",\n", c.cols, c.rows, c.pitch, c.dist, c.swz, ms, ops / (ms * 1e-3));
        first = false;
    }
    printf("\n]}\n");
    return 0;
}
