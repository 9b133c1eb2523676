// Microbenchmark: 2x2-corner gathers from an image distributed over the shared memory of a 4-CTA cluster (DSMEM)
// versus the same gathers from local shared memory and from global memory/L1.  Decides whether a cluster-distributed
// ds_dout tile is worth building for pose images that do not fit one SM's shared memory (256 x 256 Float32).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/probe_dsmem tools/probe_dsmem.cu
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float ld_cluster(uint32_t local_addr, uint32_t rank) {
    uint32_t ra; float v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra));
    return v;
}

// CSZ CTAs per cluster, each holds G/CSZ rows of a G x G image
template <int CSZ>
__global__ void __launch_bounds__(512) k_dsmem(float* sink, int G, int iters, int coherent) {
    extern __shared__ float tile[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rows = G / CSZ;
    for (int i = threadIdx.x; i < rows * G; i += blockDim.x) tile[i] = (float)(i & 7);
    cluster.sync();
    const uint32_t base = smem_u32(tile);
    float acc = 0.f;
    uint32_t seed = blockIdx.x * 1024u + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        uint32_t h = hash32(seed + it * 0x9e3779b9U);
        int x, y;
        if (coherent) { uint32_t hw = hash32((seed >> 5) + it * 0x9e3779b9U); x = (hw % (G - 17)) + (h & 15); y = ((hw >> 12) % (G - 17)) + ((h >> 4) & 15); }
        else { x = h % (G - 1); y = (h >> 12) % (G - 1); }
        const float w = (float)(h >> 24) * (1.f / 256.f);
        const int y1 = y + 1;
        const uint32_t a0 = base + (uint32_t)(((y % rows) * G + x) * 4), a1 = base + (uint32_t)(((y1 % rows) * G + x) * 4);
        const uint32_t r0 = y / rows, r1 = y1 / rows;
        acc += ld_cluster(a0, r0) * w + ld_cluster(a0 + 4, r0) * (1.f - w) + ld_cluster(a1, r1) * w + ld_cluster(a1 + 4, r1);
    }
    cluster.sync();
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

template <int CSZ>
static void run(float* sink, int G, int sms, int iters) {
    size_t sh = (size_t)(G / CSZ) * G * 4;
    CK(cudaFuncSetAttribute(k_dsmem<CSZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    int ctas = (sms / CSZ) * CSZ;
    for (int coh = 0; coh < 2; ++coh) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = sh;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = CSZ; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        double ms = time_ms([&] { CK(cudaLaunchKernelEx(&cfg, k_dsmem<CSZ>, sink, G, iters, coh)); }, 3);
        double ops = (double)ctas * 512.0 * iters * 4.0;
        printf(" {\"case\": \"dsmem_gather_cluster%d\", \"G\": %d, \"ctas\": %d, \"smem_per_cta\": %zu, \"coherent\": %d, \"ms\": %.4f, \"corner_ops_per_s\": %.4e},\n", CSZ, G, ctas, sh, coh, ms, ops / (ms * 1e-3));
        fflush(stdout);
    }
}

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    float* sink; CK(cudaMalloc(&sink, 1024));
    printf("{\"device\": \"%s\", \"sms\": %d, \"results\": [\n", p.name, p.multiProcessorCount);
    const int iters = 2048;
    run<1>(sink, 128, p.multiProcessorCount, iters);   // local shared memory only (64 KB image)
    run<2>(sink, 256, p.multiProcessorCount, iters);   // 256 x 256 over 2 CTAs (128 KB each)
    run<4>(sink, 256, p.multiProcessorCount, iters);   // over 4 CTAs (64 KB each)
    run<8>(sink, 256, p.multiProcessorCount, iters);   // over 8 CTAs (32 KB each)
    printf(" {}]}\n");
    return 0;
}
