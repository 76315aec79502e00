// Micro-benchmark: FP64 tensor-core (mma.sync m8n8k4 f64 = DMMA) peak against the DFMA peak, alone and
// mixed, to decide whether the 1-D contractions of the cell operator may use it.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_peak tools/dmma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, const double a, const double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC, int NFMA>
__global__ void k_dmma(double *out, int iters)
{
  double       c[NACC > 0 ? NACC : 1][2];
  double       f[NFMA > 0 ? NFMA : 1];
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-3 * (threadIdx.x % 7);
#pragma unroll
  for (int i = 0; i < NACC; ++i)
    c[i][0] = i, c[i][1] = -i;
#pragma unroll
  for (int i = 0; i < NFMA; ++i)
    f[i] = i + threadIdx.x;
  for (int it = 0; it < iters; ++it)
    {
#pragma unroll
      for (int i = 0; i < NACC; ++i)
        dmma(c[i][0], c[i][1], a, b);
#pragma unroll
      for (int i = 0; i < NFMA; ++i)
        f[i] = fma(f[i], a, b);
    }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i)
    s += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < NFMA; ++i)
    s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC, int NFMA>
void run(const char *name, int sms, double *out, int warps_per_sm)
{
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  const int iters = 4000, blocks = sms * warps_per_sm / 8;
  for (int rep = 0; rep < 2; ++rep)
    {
      cudaEventRecord(e0);
      k_dmma<NACC, NFMA><<<blocks, 256>>>(out, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double warps = (double)blocks * 8;
      const double mma_fma = warps * iters * NACC * 256.0, dfma = warps * 32 * iters * (double)NFMA;
      if (rep == 1)
        printf("%-28s warps/SM %2d: DMMA %.2f TFMA/s, DFMA %.2f TFMA/s (%.3f ms)\n", name, warps_per_sm, mma_fma / ms * 1e-9,
               dfma / ms * 1e-9, ms);
    }
}

int main()
{
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double   *out;
  cudaMalloc(&out, sizeof(double) * sms * 64 * 32);
  for (int w : {8, 16, 32, 64})
    {
      run<8, 0>("dmma x8 acc", sms, out, w);
      run<4, 0>("dmma x4 acc", sms, out, w);
      run<2, 0>("dmma x2 acc", sms, out, w);
      run<0, 8>("dfma x8", sms, out, w);
      run<4, 8>("dmma x4 + dfma x8", sms, out, w);
      run<4, 32>("dmma x4 + dfma x32", sms, out, w);
    }
  return 0;
}
