// Micro-benchmark: FP64 FMA peak and device-to-device copy bandwidth.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_fma(double *out, int iters)
{
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i)
    {
      a0 = fma(a0, b, c), a1 = fma(a1, b, c), a2 = fma(a2, b, c), a3 = fma(a3, b, c);
      a4 = fma(a4, b, c), a5 = fma(a5, b, c), a6 = fma(a6, b, c), a7 = fma(a7, b, c);
    }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void k_copy(double *__restrict__ d, const double *__restrict__ s, size_t n)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    d[i] = s[i];
}
int main()
{
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double   *out;
  cudaMalloc(&out, sizeof(double) * sms * 8 * 256);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  const int iters = 20000;
  for (int rep = 0; rep < 3; ++rep)
    {
      cudaEventRecord(e0);
      k_fma<<<sms * 8, 256>>>(out, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double fmas = (double)sms * 8 * 256 * iters * 8;
      printf("fp64 fma: %.2f TFMA/s = %.2f TFLOP/s (%d SMs, %.3f ms)\n", fmas / ms * 1e-9, 2 * fmas / ms * 1e-9, sms, ms);
    }
  const size_t n = (size_t)1 << 28; // 2 GiB per array
  double      *a, *b;
  cudaMalloc(&a, n * 8), cudaMalloc(&b, n * 8);
  cudaMemset(a, 0, n * 8);
  for (int rep = 0; rep < 3; ++rep)
    {
      cudaEventRecord(e0);
      k_copy<<<sms * 16, 512>>>(b, a, n);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      printf("copy: %.1f GB/s (read+write)\n", 2.0 * n * 8 / ms * 1e-6);
    }
  return 0;
}
