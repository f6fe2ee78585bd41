// tools/copybench.cu -- what HBM bandwidth does the fast kernel's ACCESS PATTERN allow, with no math?
// pattern A: every warp streams its own contiguous segment of SEG passes (512 B in + 512 B out per pass)
// pattern B: classic grid-stride (consecutive warps touch consecutive 512 B pieces)
// pattern C: a CTA of 8 warps walks one contiguous region (4 KiB per step)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) seg_copy(const uint4* __restrict__ in, float4* __restrict__ out, unsigned long long nseg, int seg, int depth) {
  const unsigned lane = threadIdx.x & 31;
  const unsigned long long w = (unsigned long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (w >= nseg) return;
  const uint4* s = in + w * seg * 32 + lane; float4* d = out + w * seg * 32 + lane;
  uint4 r0 = __ldg(s), r1 = __ldg(s + 32), r2 = __ldg(s + 64), r3;
  for (int p = 0; p < seg; p += 4) {
    if (p + 3 < seg) r3 = __ldg(s + 32 * (p + 3));
    d[32 * p] = make_float4(__uint_as_float(r0.x), __uint_as_float(r0.y), __uint_as_float(r0.z), __uint_as_float(r0.w));
    if (p + 4 < seg) r0 = __ldg(s + 32 * (p + 4));
    d[32 * (p + 1)] = make_float4(__uint_as_float(r1.x), __uint_as_float(r1.y), __uint_as_float(r1.z), __uint_as_float(r1.w));
    if (p + 5 < seg) r1 = __ldg(s + 32 * (p + 5));
    d[32 * (p + 2)] = make_float4(__uint_as_float(r2.x), __uint_as_float(r2.y), __uint_as_float(r2.z), __uint_as_float(r2.w));
    if (p + 6 < seg) r2 = __ldg(s + 32 * (p + 6));
    d[32 * (p + 3)] = make_float4(__uint_as_float(r3.x), __uint_as_float(r3.y), __uint_as_float(r3.z), __uint_as_float(r3.w));
  }
}
__global__ void __launch_bounds__(256) lin_copy(const uint4* __restrict__ in, float4* __restrict__ out, unsigned long long n) {
  unsigned long long i = (unsigned long long)blockIdx.x * 256 + threadIdx.x;
  const unsigned long long stride = (unsigned long long)gridDim.x * 256;
  for (; i + 3 * stride < n; i += 4 * stride) {
    uint4 a = __ldg(in + i), b = __ldg(in + i + stride), c = __ldg(in + i + 2 * stride), e = __ldg(in + i + 3 * stride);
    out[i] = make_float4(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), __uint_as_float(a.w));
    out[i + stride] = make_float4(__uint_as_float(b.x), __uint_as_float(b.y), __uint_as_float(b.z), __uint_as_float(b.w));
    out[i + 2 * stride] = make_float4(__uint_as_float(c.x), __uint_as_float(c.y), __uint_as_float(c.z), __uint_as_float(c.w));
    out[i + 3 * stride] = make_float4(__uint_as_float(e.x), __uint_as_float(e.y), __uint_as_float(e.z), __uint_as_float(e.w));
  }
  for (; i < n; i += stride) { uint4 a = __ldg(in + i); out[i] = make_float4(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), __uint_as_float(a.w)); }
}
// CTA walks a contiguous region: step t, warp w handles piece t*8+w
__global__ void __launch_bounds__(256) cta_copy(const uint4* __restrict__ in, float4* __restrict__ out, unsigned long long nreg, int steps) {
  if (blockIdx.x >= nreg) return;
  const uint4* s = in + (unsigned long long)blockIdx.x * steps * 256 + threadIdx.x; float4* d = out + (unsigned long long)blockIdx.x * steps * 256 + threadIdx.x;
  uint4 r0 = __ldg(s), r1 = __ldg(s + 256), r2 = __ldg(s + 512), r3;
  for (int p = 0; p < steps; p += 4) {
    if (p + 3 < steps) r3 = __ldg(s + 256 * (p + 3));
    d[256 * p] = make_float4(__uint_as_float(r0.x), __uint_as_float(r0.y), __uint_as_float(r0.z), __uint_as_float(r0.w));
    if (p + 4 < steps) r0 = __ldg(s + 256 * (p + 4));
    d[256 * (p + 1)] = make_float4(__uint_as_float(r1.x), __uint_as_float(r1.y), __uint_as_float(r1.z), __uint_as_float(r1.w));
    if (p + 5 < steps) r1 = __ldg(s + 256 * (p + 5));
    d[256 * (p + 2)] = make_float4(__uint_as_float(r2.x), __uint_as_float(r2.y), __uint_as_float(r2.z), __uint_as_float(r2.w));
    if (p + 6 < steps) r2 = __ldg(s + 256 * (p + 6));
    d[256 * (p + 3)] = make_float4(__uint_as_float(r3.x), __uint_as_float(r3.y), __uint_as_float(r3.z), __uint_as_float(r3.w));
  }
}
int main() {
  const size_t bytes = 256ull << 20;  // 256 MiB in + 256 MiB out, like one bench channel
  uint4* in; float4* out; cudaMalloc(&in, bytes); cudaMalloc(&out, bytes); cudaMemset(in, 1, bytes);
  const unsigned long long n = bytes / 16;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  auto report = [&](const char* name, float ms, int reps) { printf("%-34s %8.1f us  %7.1f GB/s\n", name, ms * 1e3 / reps, 2.0 * bytes * reps / (ms * 1e-3) / 1e9); };
  const int reps = 20; float ms;
  for (int seg : {8, 16, 32, 64, 128, 256}) {
    unsigned long long nseg = n / 32 / seg; unsigned blocks = (unsigned)((nseg + 7) / 8);
    for (int i = 0; i < 3; ++i) seg_copy<<<blocks, 256>>>(in, out, nseg, seg, 3);
    cudaEventRecord(a); for (int i = 0; i < reps; ++i) seg_copy<<<blocks, 256>>>(in, out, nseg, seg, 3); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
    char nm[64]; snprintf(nm, 64, "warp-segment copy, %d passes/seg", seg); report(nm, ms, reps);
  }
  for (int mult : {2, 4, 8, 16}) {
    unsigned blocks = 148 * mult;
    for (int i = 0; i < 3; ++i) lin_copy<<<blocks, 256>>>(in, out, n);
    cudaEventRecord(a); for (int i = 0; i < reps; ++i) lin_copy<<<blocks, 256>>>(in, out, n); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
    char nm[64]; snprintf(nm, 64, "grid-stride copy, %d CTAs/SM", mult); report(nm, ms, reps);
  }
  for (int steps : {8, 16, 64, 256}) {
    unsigned long long nreg = n / 256 / steps;
    for (int i = 0; i < 3; ++i) cta_copy<<<(unsigned)nreg, 256>>>(in, out, nreg, steps);
    cudaEventRecord(a); for (int i = 0; i < reps; ++i) cta_copy<<<(unsigned)nreg, 256>>>(in, out, nreg, steps); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
    char nm[64]; snprintf(nm, 64, "CTA-contiguous copy, %d steps", steps); report(nm, ms, reps);
  }
  for (int i = 0; i < 3; ++i) cudaMemcpyAsync(out, in, bytes, cudaMemcpyDeviceToDevice);
  cudaEventRecord(a); for (int i = 0; i < reps; ++i) cudaMemcpyAsync(out, in, bytes, cudaMemcpyDeviceToDevice); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
  report("cudaMemcpy D2D", ms, reps);
  cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
