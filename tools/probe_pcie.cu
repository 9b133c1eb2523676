// Host link probe: H2D alone, D2H alone, both at once (two streams), for one large copy and for chunked copies.
// nvcc -O2 -o tools/probe_pcie tools/probe_pcie.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
int main() {
  const size_t n = (size_t)1 << 30;
  char *h_in, *h_out, *d_in, *d_out;
  CK(cudaHostAlloc(&h_in, n, cudaHostAllocDefault)); CK(cudaHostAlloc(&h_out, n, cudaHostAllocDefault));
  memset(h_in, 1, n); memset(h_out, 2, n);
  CK(cudaMalloc(&d_in, n)); CK(cudaMalloc(&d_out, n));
  cudaStream_t s0, s1; CK(cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  auto run = [&](int mode, size_t chunk, const char* name) -> int {
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(a, s0));
      CK(cudaStreamWaitEvent(s1, a, 0));
      for (size_t o = 0; o < n; o += chunk) {
        if (mode & 1) CK(cudaMemcpyAsync(d_in + o, h_in + o, chunk, cudaMemcpyHostToDevice, s0));
        if (mode & 2) CK(cudaMemcpyAsync(h_out + o, d_out + o, chunk, cudaMemcpyDeviceToHost, s1));
      }
      CK(cudaEventRecord(b, s1)); CK(cudaStreamWaitEvent(s0, b, 0));
      CK(cudaEventRecord(b, s0));
      CK(cudaEventSynchronize(b));
      float ms; CK(cudaEventElapsedTime(&ms, a, b));
      if (rep == 2) printf("%-28s chunk %5zu MB: %7.2f ms  %6.1f GB/s total\n", name, chunk >> 20, ms, ((mode & 1 ? n : 0) + (mode & 2 ? n : 0)) / ms * 1e-6);
    }
    return 0;
  };
  run(1, n, "H2D alone"); run(2, n, "D2H alone"); run(3, n, "H2D + D2H");
  run(3, (size_t)64 << 20, "H2D + D2H"); run(3, (size_t)8 << 20, "H2D + D2H"); run(1, (size_t)8 << 20, "H2D alone"); run(2, (size_t)8 << 20, "D2H alone");
  // pageable-registered memory (cudaHostRegister), as a caller's Array would be after registration
  char* p = (char*)aligned_alloc(4096, n); memset(p, 3, n);
  CK(cudaHostRegister(p, n, cudaHostRegisterDefault));
  char* keep = h_out; h_out = p; run(2, n, "D2H alone (registered)"); run(3, n, "H2D + D2H (registered)"); h_out = keep;
  return 0;
}
