// probe_tma_tensor.cu - does tensor-map TMA (cp.async.bulk.tensor, SASS UTMALDG) run on this pool's B200s?
//
// Round 1 declared it unusable ("illegal instruction") without a committed log.  This probe
//   1. prints the driver / runtime versions and how cuTensorMapEncodeTiled was resolved (dlsym on libcuda.so.1 AND
//      cudaGetDriverEntryPoint), and compares the two 128-byte descriptors;
//   2. runs 2-d, 3-d and 4-d tiled loads written in raw PTX (box partly outside the tensor: zero fill) and checks every
//      value;
//   3. times 4-d box loads of (32, 16, 16) Float32 tiles over a (256, 256, 256, 16) tensor - the access pattern of the
//      3-d tile pullback - with the tile summed from shared memory (d_background's pattern), in GB/s.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/probe_tma_tensor tools/probe_tma_tensor.cu -ldl
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra D_%=;\n\t"
        "bra W_%=;\n\t"
        "D_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}

template <int RANK>
__global__ void check_kernel(const __grid_constant__ CUtensorMap map, float* out, int n, int c0, int c1, int c2, int c3) {
    extern __shared__ __align__(128) unsigned char raw[];
    float* tile = reinterpret_cast<float*>(raw);
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, (uint32_t)n * 4u);
        if (RANK == 2) tma_load_2d(tile, &map, c0, c1, &bar);
        if (RANK == 3) tma_load_3d(tile, &map, c0, c1, c2, &bar);
        if (RANK == 4) tma_load_4d(tile, &map, c0, c1, c2, c3, &bar);
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = tile[i];
}

// persistent CTAs, STAGES-deep ring of (TX, TY, TZ) tiles, the tile is summed from shared memory by all threads
template <int STAGES>
__global__ void __launch_bounds__(256) stream_kernel(const __grid_constant__ CUtensorMap map, float* sums, int ntx, int nty, int ntz,
                                                     int B, int TX, int TY, int TZ) {
    extern __shared__ __align__(128) unsigned char raw[];
    float* tiles = reinterpret_cast<float*>(raw);
    __shared__ __align__(8) uint64_t full[STAGES];
    const int n = TX * TY * TZ;
    const int total = ntx * nty * ntz * B;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int t, int s) {
        const int tx = t % ntx, ty = (t / ntx) % nty, tz = (t / (ntx * nty)) % ntz, b = t / (ntx * nty * ntz);
        mbar_expect_tx(&full[s], (uint32_t)n * 4u);
        tma_load_4d(tiles + (size_t)s * n, &map, tx * TX, ty * TY, tz * TZ, b, &full[s]);
    };
    int t_issue = blockIdx.x;
    if (threadIdx.x == 0)
        for (int s = 0; s < STAGES - 1 && t_issue < total; ++s, t_issue += gridDim.x) issue(t_issue, s);
    float acc = 0.f;
    int it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        const int s = it % STAGES;
        if (threadIdx.x == 0) {
            const int tn = t + (STAGES - 1) * gridDim.x;
            if (tn < total) issue(tn, (it + STAGES - 1) % STAGES);
        }
        mbar_wait(&full[s], (it / STAGES) & 1);
        const float4* v = reinterpret_cast<const float4*>(tiles + (size_t)s * n);
        for (int i = threadIdx.x; i < n / 4; i += blockDim.x) {
            const float4 q = v[i];
            acc += (q.x + q.y) + (q.z + q.w);
        }
        __syncthreads();   // everyone is done with stage s before it is refilled
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(sums + blockIdx.x % 64, acc);
}

static void dump(const char* name, const CUtensorMap& m) {
    const unsigned* w = reinterpret_cast<const unsigned*>(&m);
    printf("%s:", name);
    for (int i = 0; i < 32; ++i) printf(" %08x", w[i]);
    printf("\n");
}

int main() {
    int drv = 0, rt = 0;
    cudaDriverGetVersion(&drv);
    cudaRuntimeGetVersion(&rt);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    printf("device %s sm_%d%d, driver API %d, runtime %d\n", prop.name, prop.major, prop.minor, drv, rt);

    EncodeFn enc_dl = nullptr, enc_ep = nullptr;
    void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (h) enc_dl = (EncodeFn)dlsym(h, "cuTensorMapEncodeTiled");
    {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        printf("cudaGetDriverEntryPoint: %s, query %d, fn %p; dlsym fn %p\n", cudaGetErrorString(e), (int)q, fn, (void*)enc_dl);
        enc_ep = (EncodeFn)fn;
    }
    EncodeFn enc = enc_dl ? enc_dl : enc_ep;
    if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 2; }

    int fails = 0;
    // ---- correctness: 2-d / 3-d / 4-d boxes that hang over the edges ----------------------------------
    {
        const int g0 = 192, g1 = 40, g2 = 12, B = 3;
        std::vector<float> hsrc((size_t)g0 * g1 * g2 * B);
        for (size_t i = 0; i < hsrc.size(); ++i) hsrc[i] = (float)(i % 100003) + 1.f;
        float* src;
        cudaMalloc(&src, hsrc.size() * 4);
        cudaMemcpy(src, hsrc.data(), hsrc.size() * 4, cudaMemcpyHostToDevice);
        for (int rank = 2; rank <= 4; ++rank) {
            cuuint64_t dims[4] = {(cuuint64_t)g0, (cuuint64_t)g1, (cuuint64_t)g2, (cuuint64_t)B};
            if (rank == 2) dims[1] = (cuuint64_t)g1 * g2 * B;
            if (rank == 3) dims[2] = (cuuint64_t)g2 * B;
            cuuint64_t strides[3] = {(cuuint64_t)g0 * 4, (cuuint64_t)g0 * g1 * 4, (cuuint64_t)g0 * g1 * g2 * 4};
            cuuint32_t box[4] = {36, 9, (cuuint32_t)(rank >= 3 ? 5 : 1), 1};
            cuuint32_t es[4] = {1, 1, 1, 1};
            CUtensorMap map, map2;
            memset(&map, 0, sizeof(map));
            memset(&map2, 0, sizeof(map2));
            CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            printf("rank %d: encode rc=%d\n", rank, (int)r);
            if (enc_ep && enc_dl) {
                enc_ep(&map2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                printf("rank %d: descriptors from dlsym and entry point %s\n", rank, memcmp(&map, &map2, sizeof(map)) ? "DIFFER" : "identical");
            }
            if (rank == 4) dump("map4d", map);
            const int n = box[0] * box[1] * box[2];
            float* out;
            cudaMalloc(&out, n * 4);
            const int c0 = -3, c1 = g1 - 4, c2 = (rank >= 3) ? -2 : 0, c3 = 1;
            if (rank == 2) check_kernel<2><<<1, 128, n * 4>>>(map, out, n, c0, c1, 0, 0);
            if (rank == 3) check_kernel<3><<<1, 128, n * 4>>>(map, out, n, c0, c1, c2, 0);
            if (rank == 4) check_kernel<4><<<1, 128, n * 4>>>(map, out, n, c0, c1, c2, c3);
            cudaError_t e = cudaDeviceSynchronize();
            printf("rank %d: kernel: %s\n", rank, cudaGetErrorString(e));
            if (e != cudaSuccess) return 1;
            std::vector<float> o(n);
            cudaMemcpy(o.data(), out, n * 4, cudaMemcpyDeviceToHost);
            size_t bad = 0;
            for (int z = 0; z < (int)box[2]; ++z)
                for (int y = 0; y < (int)box[1]; ++y)
                    for (int x = 0; x < (int)box[0]; ++x) {
                        const long gx = c0 + x, gy = c1 + y, gz = c2 + z;
                        float want = 0.f;
                        if (rank == 2) {
                            if (gx >= 0 && gx < g0 && gy >= 0 && gy < (long)g1 * g2 * B) want = hsrc[(size_t)gy * g0 + gx];
                        } else if (rank == 3) {
                            if (gx >= 0 && gx < g0 && gy >= 0 && gy < g1 && gz >= 0 && gz < (long)g2 * B) want = hsrc[((size_t)gz * g1 + gy) * g0 + gx];
                        } else {
                            if (gx >= 0 && gx < g0 && gy >= 0 && gy < g1 && gz >= 0 && gz < g2)
                                want = hsrc[(((size_t)c3 * g2 + gz) * g1 + gy) * g0 + gx];
                        }
                        if (o[((size_t)z * box[1] + y) * box[0] + x] != want) ++bad;
                    }
            printf("rank %d: mismatches %zu of %d\n", rank, bad, n);
            if (bad) ++fails;
            cudaFree(out);
        }
        cudaFree(src);
    }
    // ---- bandwidth: (32,16,16) and (32,32,16) tiles of a (256,256,256,16) Float32 tensor ---------------
    {
        const int g = 256, B = 16;
        const size_t total = (size_t)g * g * g * B;
        float* src;
        if (cudaMalloc(&src, total * 4) != cudaSuccess) { printf("cudaMalloc failed\n"); return 3; }
        cudaMemset(src, 0, total * 4);
        float* sums;
        cudaMalloc(&sums, 64 * 4);
        cudaMemset(sums, 0, 64 * 4);
        const int shapes[3][3] = {{32, 16, 16}, {32, 32, 16}, {64, 16, 16}};
        for (int sh = 0; sh < 3; ++sh) {
            const int TX = shapes[sh][0], TY = shapes[sh][1], TZ = shapes[sh][2];
            cuuint64_t dims[4] = {g, g, g, B};
            cuuint64_t strides[3] = {(cuuint64_t)g * 4, (cuuint64_t)g * g * 4, (cuuint64_t)g * g * g * 4};
            cuuint32_t box[4] = {(cuuint32_t)TX, (cuuint32_t)TY, (cuuint32_t)TZ, 1};
            cuuint32_t es[4] = {1, 1, 1, 1};
            CUtensorMap map;
            CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 4; }
            const int tile_bytes = TX * TY * TZ * 4;
            for (int ctas_per_sm = 1; ctas_per_sm <= 4; ctas_per_sm *= 2) {
                const int stages = 3;
                const int smem = stages * tile_bytes;
                if ((size_t)smem * ctas_per_sm > 220 * 1024) continue;
                cudaFuncSetAttribute(stream_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                cudaEvent_t e0, e1;
                cudaEventCreate(&e0);
                cudaEventCreate(&e1);
                const int grid = prop.multiProcessorCount * ctas_per_sm;
                for (int rep = 0; rep < 2; ++rep) {
                    cudaEventRecord(e0);
                    stream_kernel<3><<<grid, 256, smem>>>(map, sums, g / TX, g / TY, g / TZ, B, TX, TY, TZ);
                    cudaEventRecord(e1);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("stream kernel: %s\n", cudaGetErrorString(e)); return 1; }
                }
                float ms = 0.f;
                cudaEventElapsedTime(&ms, e0, e1);
                printf("{\"probe\": \"tma4d_stream\", \"tile\": [%d, %d, %d], \"ctas_per_sm\": %d, \"stages\": %d, \"ms\": %.4f, \"GBps\": %.1f}\n", TX,
                       TY, TZ, ctas_per_sm, stages, ms, total * 4.0 / (ms * 1e-3) / 1e9);
            }
        }
        cudaFree(src);
    }
    printf(fails ? "PROBE FAILED\n" : "PROBE OK\n");
    return fails ? 1 : 0;
}
