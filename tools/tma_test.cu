// Stand-alone check of the 2-D FP64 tensor-map staging used by op_v3.cuh (super-row view of an odd-pitch array).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_test tools/tma_test.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

typedef CUresult (*PFN_enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                            const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
struct Args
{
  int    c0, c1, pad0, pad1;
  double *out;
  alignas(64) CUtensorMap tm;
};

template <int BW, int BH>
__global__ void k(const __grid_constant__ Args a)
{
  extern __shared__ __align__(16) double raw[];
  double   *sm  = reinterpret_cast<double *>((reinterpret_cast<uintptr_t>(raw) + 127) & ~(uintptr_t)127);
  uint64_t *bar = reinterpret_cast<uint64_t *>(sm + BW * BH + 16);
  const unsigned b32 = (unsigned)__cvta_generic_to_shared(bar), d32 = (unsigned)__cvta_generic_to_shared(sm);
  if (threadIdx.x == 0)
    {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(b32) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
  __syncthreads();
  if (threadIdx.x == 0)
    {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(b32), "r"(BW * BH * 8) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(d32),
                   "l"(reinterpret_cast<uint64_t>(&a.tm)), "r"(a.c0), "r"(a.c1), "r"(b32)
                   : "memory");
    }
  unsigned ok = 0;
  while (!ok)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(b32), "r"(0) : "memory");
  for (int i = threadIdx.x; i < BW * BH; i += blockDim.x)
    a.out[i] = sm[i];
}

int main(int argc, char **argv)
{
  const int n1 = 33, shift = argc > 1 ? atoi(argv[1]) : 0, l2 = argc > 2 ? atoi(argv[2]) : 3;
  constexpr int BW = 38, BH = 19;
  const long long N = (long long)n1 * n1 * n1;
  std::vector<double> h(N + 2);
  for (long long i = 0; i < N + 2; ++i)
    h[i] = (double)i;
  double *d, *out;
  cudaMalloc(&d, (N + 2) * 8), cudaMalloc(&out, BW * BH * 8);
  cudaMemcpy(d, h.data(), (N + 2) * 8, cudaMemcpyHostToDevice);
  void *fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  PFN_enc enc = (PFN_enc)fp;
  Args a;
  const double *ptr = d + shift; // vector starts at d + shift (8-byte aligned only if shift odd)
  const uintptr_t p0 = (uintptr_t)ptr;
  const int sh = (int)((p0 >> 3) & 1);
  const cuuint64_t dims[2] = {(cuuint64_t)2 * n1, (cuuint64_t)((N + sh) / (2 * n1))}, strides[1] = {(cuuint64_t)2 * n1 * 8};
  const cuuint32_t box[2] = {BW, BH}, es[2] = {1, 1};
  CUresult r = enc(&a.tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)(p0 & ~(uintptr_t)15), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d (shift %d, sh %d)\n", (int)r, shift, sh);
  // staged rows of plane P=2, gy0-K = -4 ... : Rb = n1*2 - 4; even box
  const long long Rb = (long long)n1 * 2 - 4;
  const int odd = argc > 3 ? atoi(argv[3]) : 0; // 0: even-row box, 1: odd-row box
  const int cw = (odd ? n1 : 0) - 4 + sh; // wanted first element
  const int exact = argc > 4 ? atoi(argv[4]) : 0; // 1: the box starts exactly at the wanted element (odd coordinates allowed?)
  a.c0 = exact ? cw : (cw & ~1), a.c1 = odd ? (int)(Rb >> 1) : (int)((Rb + 1) >> 1), a.out = out;
  printf("c0 = %d (exact %d)\n", a.c0, exact);
  k<BW, BH><<<1, 128, (BW * BH + 64) * 8 + 128>>>(a);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess)
    return 1;
  std::vector<double> o(BW * BH);
  cudaMemcpy(o.data(), out, BW * BH * 8, cudaMemcpyDeviceToHost);
  // expected: row i of the box = node row R = 2*(c1+i) (even rows), x = -4 .. 33: value = index relative to ptr
  int bad = 0;
  for (int i = 0; i < BH; ++i)
    for (int j = 0; j < BW; ++j)
      {
        // element (i, j) of the box = element a.c0 + j of super-row a.c1 + i of the map = index relative to the map base
        const long long e_ = ((long long)a.c1 + i) * 2 * n1 + a.c0 + j;
        double expect = (a.c0 + j < 0 || a.c0 + j >= 2 * n1) ? 0.0 : (double)(e_ + shift - sh);
        if (o[i * BW + j] != expect && bad++ < 5)
          printf(" mismatch box(%d,%d): got %.0f expect %.0f\n", i, j, o[i * BW + j], expect);
      }
  printf("mismatches: %d\n", bad);
  return 0;
}
