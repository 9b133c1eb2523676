// Microbenchmark: on-chip ceilings that bound the splat (scatter) and pullback (gather) stencils on B200.
// Measures lane-operations per second for: shared-memory float atomicAdd (CAS loop on sm_100a),
// shared-memory int32 atomicAdd (native ATOMS.ADD), global REDG f32 / f32x2 into an L2-resident
// per-CTA image, LDS gathers and LDG gathers, each with the 2-d 4-corner stencil access pattern.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/probe_atomics tools/probe_atomics.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

// G = image edge (G x G floats). mode: 0 smem f32 CAS, 1 smem i32 native, 2 LDS gather
template <int MODE>
__global__ void __launch_bounds__(1024) k_smem(float* sink, int G, int iters, int coherent) {
    extern __shared__ float tile[];
    int n = G * G;
    for (int i = threadIdx.x; i < n; i += blockDim.x) tile[i] = 0.f;
    __syncthreads();
    float acc = 0.f;
    uint32_t seed = blockIdx.x * 1024u + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        uint32_t h = hash32(seed + it * 0x9e3779b9U);
        int x, y;
        if (coherent) { // neighbouring lanes land in a small neighbourhood (sorted points)
            uint32_t hw = hash32((seed >> 5) + it * 0x9e3779b9U);
            x = (hw % (G - 9)) + (h & 7); y = ((hw >> 12) % (G - 9)) + ((h >> 3) & 7);
        } else { x = h % (G - 1); y = (h >> 12) % (G - 1); }
        int base = y * G + x;
        float w = (float)(h >> 24) * (1.f / 256.f);
        if (MODE == 0) {
            atomicAdd(&tile[base], w); atomicAdd(&tile[base + 1], 1.f - w);
            atomicAdd(&tile[base + G], w * 0.5f); atomicAdd(&tile[base + G + 1], 0.5f - w * 0.5f);
        } else if (MODE == 1) {
            int* ti = reinterpret_cast<int*>(tile); int q = (int)(w * 4096.f);
            atomicAdd(&ti[base], q); atomicAdd(&ti[base + 1], 4096 - q);
            atomicAdd(&ti[base + G], q >> 1); atomicAdd(&ti[base + G + 1], 2048 - (q >> 1));
        } else {
            acc += tile[base] * w + tile[base + 1] * (1.f - w) + tile[base + G] * w + tile[base + G + 1];
        }
    }
    __syncthreads();
    float s = acc;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += tile[i];
    if (s == 123.456f) sink[0] = s;
}

// mode: 0 REDG f32 scalar x4, 1 REDG f32x2 when aligned else scalar, 2 LDG gather, 3 REDG v2 always aligned (x even)
template <int MODE>
__global__ void __launch_bounds__(1024) k_gmem(float* img, float* sink, int G, int iters, int coherent) {
    float* my = img + (size_t)blockIdx.x * G * G;
    float acc = 0.f;
    uint32_t seed = blockIdx.x * 1024u + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        uint32_t h = hash32(seed + it * 0x9e3779b9U);
        int x, y;
        if (coherent) {
            uint32_t hw = hash32((seed >> 5) + it * 0x9e3779b9U);
            x = (hw % (G - 9)) + (h & 7); y = ((hw >> 12) % (G - 9)) + ((h >> 3) & 7);
        } else { x = h % (G - 1); y = (h >> 12) % (G - 1); }
        if (MODE == 3) x &= ~1;
        int base = y * G + x;
        float w = (float)(h >> 24) * (1.f / 256.f);
        if (MODE == 0) {
            atomicAdd(&my[base], w); atomicAdd(&my[base + 1], 1.f - w);
            atomicAdd(&my[base + G], w * 0.5f); atomicAdd(&my[base + G + 1], 0.5f - w * 0.5f);
        } else if (MODE == 1 || MODE == 3) {
            if ((base & 1) == 0) {
                atomicAdd(reinterpret_cast<float2*>(&my[base]), make_float2(w, 1.f - w));
                atomicAdd(reinterpret_cast<float2*>(&my[base + G]), make_float2(w * 0.5f, 0.5f - w * 0.5f));
            } else {
                atomicAdd(&my[base], w); atomicAdd(&my[base + 1], 1.f - w);
                atomicAdd(&my[base + G], w * 0.5f); atomicAdd(&my[base + G + 1], 0.5f - w * 0.5f);
            }
        } else {
            acc += __ldg(&my[base]) * w + __ldg(&my[base + 1]) * (1.f - w) + __ldg(&my[base + G]) * w + __ldg(&my[base + G + 1]);
        }
    }
    if (acc == 123.456f) sink[0] = acc;
}

template <typename F>
static double time_ms(F launch, int reps) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    launch(); launch(); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) launch();
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    CK(cudaGetLastError());
    return ms / reps;
}

int main() {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    int sms = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"results\": [\n", p.name, sms);
    float* sink; CK(cudaMalloc(&sink, 1024));
    const int iters = 2048;
    bool first = true;
    auto report = [&](const char* name, int G, int ctas, int coherent, double ms) {
        double ops = (double)ctas * 1024.0 * iters * 4.0;
        printf("%s {\"case\": \"%s\", \"G\": %d, \"ctas\": %d, \"coherent\": %d, \"ms\": %.4f, \"corner_ops_per_s\": %.4e}",
               first ? "" : ",\n", name, G, ctas, coherent, ms, ops / (ms * 1e-3));
        first = false; fflush(stdout);
    };
    CK(cudaFuncSetAttribute(k_smem<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_smem<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_smem<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int Gs[2] = {128, 220};
    for (int gi = 0; gi < 2; ++gi) {
        int G = Gs[gi]; size_t sh = (size_t)G * G * 4;
        int per_sm = (sh <= 100 * 1024) ? 2 : 1;
        int ctas = sms * per_sm * 2;
        for (int coh = 0; coh < 2; ++coh) {
            report("smem_f32_cas", G, ctas, coh, time_ms([&] { k_smem<0><<<ctas, 1024, sh>>>(sink, G, iters, coh); }, 5));
            report("smem_i32_native", G, ctas, coh, time_ms([&] { k_smem<1><<<ctas, 1024, sh>>>(sink, G, iters, coh); }, 5));
            report("smem_lds_gather", G, ctas, coh, time_ms([&] { k_smem<2><<<ctas, 1024, sh>>>(sink, G, iters, coh); }, 5));
        }
    }
    int Gg[3] = {128, 256, 512};
    for (int gi = 0; gi < 3; ++gi) {
        int G = Gg[gi];
        int ctas = sms * 2;
        float* img; CK(cudaMalloc(&img, (size_t)ctas * G * G * 4)); CK(cudaMemset(img, 0, (size_t)ctas * G * G * 4));
        for (int coh = 0; coh < 2; ++coh) {
            report("gmem_redg_f32", G, ctas, coh, time_ms([&] { k_gmem<0><<<ctas, 1024>>>(img, sink, G, iters, coh); }, 3));
            report("gmem_redg_f32x2_mixed", G, ctas, coh, time_ms([&] { k_gmem<1><<<ctas, 1024>>>(img, sink, G, iters, coh); }, 3));
            report("gmem_redg_f32x2_aligned", G, ctas, coh, time_ms([&] { k_gmem<3><<<ctas, 1024>>>(img, sink, G, iters, coh); }, 3));
            report("gmem_ldg_gather", G, ctas, coh, time_ms([&] { k_gmem<2><<<ctas, 1024>>>(img, sink, G, iters, coh); }, 3));
        }
        CK(cudaFree(img));
    }
    printf("\n]}\n");
    return 0;
}
