// Scratch probe 2: the CUDA programming guide's own TMA example (libcu++ wrappers), 2-d and 3-d.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
constexpr int WX = 64, WY = 64;
template <int RANK>
__global__ void k(const __grid_constant__ CUtensorMap tmap, float* out, int c0, int c1, int c2) {
    __shared__ alignas(128) float tile[WY][WX];
    #pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        if (RANK == 2) cde::cp_async_bulk_tensor_2d_global_to_shared(&tile, &tmap, c0, c1, bar);
        else cde::cp_async_bulk_tensor_3d_global_to_shared(&tile, &tmap, c0, c1, c2, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(tile));
    } else token = bar.arrive();
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < WX * WY; i += blockDim.x) out[i] = (&tile[0][0])[i];
}
int main(int argc, char** argv) {
    const int rank = argc > 1 ? atoi(argv[1]) : 2;
    void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    Fn enc = (Fn)fnp;
    const int g0 = 192, g1 = 160, B = 9;
    std::vector<float> h((size_t)g0 * g1 * B);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100003) + 1.f;
    float* img; cudaMalloc(&img, h.size() * 4); cudaMemcpy(img, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    float* out; cudaMalloc(&out, WX * WY * 4);
    CUtensorMap map;
    cuuint64_t dims[3] = {(cuuint64_t)g0, (cuuint64_t)(rank == 2 ? g1 * B : g1), (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)g0 * 4, (cuuint64_t)g0 * g1 * 4};
    cuuint32_t box[3] = {WX, WY, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, img, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("rank %d encode rc=%d\n", rank, (int)r);
    const int c0 = -10, c1 = 20, c2 = 3;
    if (rank == 2) k<2><<<1, 128>>>(map, out, c0, c1, c2); else k<3><<<1, 128>>>(map, out, c0, c1, c2);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> o(WX * WY); cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
    size_t bad = 0;
    for (int y = 0; y < WY; ++y) for (int x = 0; x < WX; ++x) {
        const int gx = c0 + x, gy = c1 + y;
        const size_t base = rank == 2 ? 0 : (size_t)c2 * g0 * g1;
        const float want = (gx >= 0 && gx < g0 && gy >= 0 && gy < (rank == 2 ? g1 * B : g1)) ? h[base + (size_t)gy * g0 + gx] : 0.f;
        if (o[(size_t)y * WX + x] != want) ++bad;
    }
    printf("mismatches %zu\n", bad);
    return 0;
}
