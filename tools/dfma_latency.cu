// Micro-benchmark: dependent-issue latency of DFMA (register, uniform-register and constant operands) measured
// with clock64 on one warp, and the chains needed per warp to fill the FP64 pipe.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dfma_latency tools/dfma_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double c_k[4] = {1.0000001, 0.9999999, 1e-9, -1e-9};

template <int NCH>
__global__ void k_lat(double *out, long long *cyc, int iters, double bb)
{
  double a[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i)
    a[i] = threadIdx.x * 1e-3 + i;
  const double b = bb, c = c_k[2];
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it)
    {
#pragma unroll
      for (int i = 0; i < NCH; ++i)
        a[i] = fma(a[i], b, c);
    }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NCH; ++i)
    s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0)
    *cyc = t1 - t0;
}
template <int NCH>
void run(double *out, long long *cyc, int warps)
{
  const int iters = 4096;
  k_lat<NCH><<<1, 32 * warps>>>(out, cyc, iters, 1.0000001);
  k_lat<NCH><<<1, 32 * warps>>>(out, cyc, iters, 1.0000001);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("warps/SM %2d chains/thread %2d: %.2f cycles per DFMA per warp, %.2f cycles per loop iteration\n", warps, NCH,
         (double)h / iters / NCH, (double)h / iters);
}
int main()
{
  double *out;
  long long *cyc;
  cudaMalloc(&out, 8 * 1024), cudaMalloc(&cyc, 8);
  for (int w : {1, 4, 8, 16})
    {
      run<1>(out, cyc, w);
      run<2>(out, cyc, w);
      run<4>(out, cyc, w);
      run<8>(out, cyc, w);
      run<16>(out, cyc, w);
    }
  return 0;
}
