// Scratch probe: 3-d tensor-map TMA box loads (negative / out-of-range coordinates, zero fill) for several box shapes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/probe_tma_box tools/probe_tma_box.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1;} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k_box(const __grid_constant__ CUtensorMap tmap, float* out, int wx, int wy, int c0, int c1, int c2) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    float* tile = reinterpret_cast<float*>(smem);
    const uint32_t bytes = (uint32_t)wx * wy * 4;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smem_u32(tile)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
    for (int i = threadIdx.x; i < wx * wy; i += blockDim.x) out[i] = tile[i];
}
__global__ void k_box_g(const CUtensorMap* gmap, float* out, int wx, int wy, int c0, int c1, int c2) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    float* tile = reinterpret_cast<float*>(smem);
    const uint32_t bytes = (uint32_t)wx * wy * 4;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smem_u32(tile)), "l"(reinterpret_cast<uint64_t>(gmap)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tW2: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D2;\n\tbra W2;\n\tD2:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
    for (int i = threadIdx.x; i < wx * wy; i += blockDim.x) out[i] = tile[i];
}
int main(int argc, char** argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    CK(cudaSetDevice(0));
    void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    Fn enc = (Fn)fnp;
    const int g0 = 192, g1 = 160, B = 9;
    std::vector<float> h((size_t)g0 * g1 * B);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100003) + 1.f;
    float* img; CK(cudaMalloc(&img, h.size() * 4)); CK(cudaMemcpy(img, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    float* out; CK(cudaMalloc(&out, 256 * 256 * 4));
    struct Case { int wx, wy, c0, c1, c2; };
    const Case cases[] = {{64, 64, 10, 20, 3}, {128, 128, 10, 20, 3}, {136, 128, 10, 20, 3}, {136, 128, -30, -40, 8}, {136, 128, 150, 100, 0}, {168, 160, -5, 60, 2}, {256, 64, 0, 0, 1}};
    for (const Case& c : cases) {
        CUtensorMap map;
        cuuint64_t dims[3] = {(cuuint64_t)g0, (cuuint64_t)g1, (cuuint64_t)B};
        cuuint64_t strides[2] = {(cuuint64_t)g0 * 4, (cuuint64_t)g0 * g1 * 4};
        cuuint32_t box[3] = {(cuuint32_t)c.wx, (cuuint32_t)c.wy, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, img, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("mode %d box %dx%d at (%d,%d,%d): encode rc=%d ", mode, c.wx, c.wy, c.c0, c.c1, c.c2, (int)r);
        { const unsigned long long* w = reinterpret_cast<const unsigned long long*>(&map); printf("[desc %016llx %016llx %016llx %016llx] ", w[0], w[1], w[2], w[3]); }
        if (r != CUDA_SUCCESS) { printf("\n"); continue; }
        const size_t smem = (size_t)c.wx * c.wy * 4;
        CK(cudaFuncSetAttribute(k_box, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (mode == 0) k_box<<<1, 256, smem>>>(map, out, c.wx, c.wy, c.c0, c.c1, c.c2);
        else {
            CUtensorMap* gmap; CK(cudaMalloc(&gmap, sizeof(CUtensorMap))); CK(cudaMemcpy(gmap, &map, sizeof(CUtensorMap), cudaMemcpyHostToDevice));
            CK(cudaFuncSetAttribute(k_box_g, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_box_g<<<1, 256, smem>>>(gmap, out, c.wx, c.wy, c.c0, c.c1, c.c2);
        }
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel error: %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<float> o((size_t)c.wx * c.wy);
        CK(cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost));
        size_t bad = 0;
        for (int y = 0; y < c.wy; ++y) for (int x = 0; x < c.wx; ++x) {
            const int gx = c.c0 + x, gy = c.c1 + y;
            const float want = (gx >= 0 && gx < g0 && gy >= 0 && gy < g1) ? h[(size_t)c.c2 * g0 * g1 + (size_t)gy * g0 + gx] : 0.f;
            if (o[(size_t)y * c.wx + x] != want) ++bad;
        }
        printf("mismatches %zu\n", bad);
    }
    return 0;
}
