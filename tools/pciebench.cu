// tools/pciebench.cu -- host<->device copy ceilings on this box (pinned memory), for the e2e number.
#include <cstdio>
#include <cuda_runtime.h>
int main() {
  const size_t n = 512ull << 20; void *h1, *h2, *d1, *d2;
  cudaHostAlloc(&h1, n, cudaHostAllocPortable); cudaHostAlloc(&h2, n, cudaHostAllocPortable); cudaMalloc(&d1, n); cudaMalloc(&d2, n);
  cudaStream_t a, b; cudaStreamCreate(&a); cudaStreamCreate(&b);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0, a); for (int i = 0; i < 4; ++i) cudaMemcpyAsync(d1, h1, n, cudaMemcpyHostToDevice, a); cudaEventRecord(e1, a); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    if (rep) printf("H2D alone        %.1f GB/s\n", 4.0 * n / ms / 1e6);
    cudaEventRecord(e0, a); for (int i = 0; i < 4; ++i) cudaMemcpyAsync(h2, d2, n, cudaMemcpyDeviceToHost, a); cudaEventRecord(e1, a); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    if (rep) printf("D2H alone        %.1f GB/s\n", 4.0 * n / ms / 1e6);
    cudaDeviceSynchronize();
    cudaEventRecord(e0, a); for (int i = 0; i < 4; ++i) { cudaMemcpyAsync(d1, h1, n, cudaMemcpyHostToDevice, a); cudaMemcpyAsync(h2, d2, n, cudaMemcpyDeviceToHost, b); }
    cudaStreamSynchronize(b); cudaEventRecord(e1, a); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    if (rep) printf("H2D + D2H both   %.1f GB/s each direction (%.1f total)\n", 4.0 * n / ms / 1e6, 8.0 * n / ms / 1e6);
  }
  { // write-combined input buffer: does the GPU read it faster?
    void* hw; cudaHostAlloc(&hw, n, cudaHostAllocPortable | cudaHostAllocWriteCombined);
    for (int rep = 0; rep < 2; ++rep) {
      cudaDeviceSynchronize(); cudaEventRecord(e0, a);
      for (int i = 0; i < 4; ++i) cudaMemcpyAsync(d1, hw, n, cudaMemcpyHostToDevice, a);
      cudaEventRecord(e1, a); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
      if (rep) printf("H2D alone, write-combined source  %.1f GB/s\n", 4.0 * n / ms / 1e6);
      cudaDeviceSynchronize(); cudaEventRecord(e0, a);
      for (int i = 0; i < 4; ++i) { cudaMemcpyAsync(d1, hw, n, cudaMemcpyHostToDevice, a); cudaMemcpyAsync(h2, d2, n, cudaMemcpyDeviceToHost, b); }
      cudaStreamSynchronize(b); cudaEventRecord(e1, a); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
      if (rep) printf("H2D (WC source) + D2H both        %.1f GB/s each direction\n", 4.0 * n / ms / 1e6);
    }
  }
  // slice size sensitivity, both directions
  for (size_t s : {4ull << 20, 16ull << 20, 32ull << 20, 64ull << 20, 128ull << 20}) {
    cudaDeviceSynchronize(); cudaEventRecord(e0, a);
    for (size_t off = 0; off < n; off += s) { cudaMemcpyAsync((char*)d1 + off, (char*)h1 + off, s, cudaMemcpyHostToDevice, a); cudaMemcpyAsync((char*)h2 + off, (char*)d2 + off, s, cudaMemcpyDeviceToHost, b); }
    cudaStreamSynchronize(b); cudaEventRecord(e1, a); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("both, %3zu MiB slices  %.1f GB/s each direction\n", s >> 20, (double)n / ms / 1e6);
  }
  return 0;
}
