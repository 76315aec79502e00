// Shared declarations of the CUDA device layer (implementation of include/spirk_b200.h).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/spirk_b200.h"

#define SPIRK_MAX_DEGREE 6
#define SPIRK_MAX_N (SPIRK_MAX_DEGREE + 1)

namespace spirk
{
  // per-degree reference 1-D matrices in constant memory (filled once per process)
  struct FeConst
  {
    double Mh[SPIRK_MAX_N * SPIRK_MAX_N];                 // [i*n+j]
    double Kh[SPIRK_MAX_N * SPIRK_MAX_N];
    double P[(2 * SPIRK_MAX_DEGREE + 1) * SPIRK_MAX_N];   // [(r)*n+i]
    double Be[(SPIRK_MAX_DEGREE + 2) * SPIRK_MAX_N];      // [(q)*n+i]
    double xe[SPIRK_MAX_DEGREE + 2], we[SPIRK_MAX_DEGREE + 2];
    double nodes[SPIRK_MAX_N];
    double Mv, Kv; // assembled 1-D diagonals at a vertex node shared by two cells: Mh[k][k] + Mh[0][0]
  };

  // geometry handed to kernels by value
  struct Geo
  {
    int       dim, k, n, nc, n1;
    long long N;
    double    h;
  };

  inline Geo make_geo(const spirk_level *l)
  {
    Geo g;
    g.dim = l->dim, g.k = l->degree, g.n = l->degree + 1, g.nc = l->n_cells_1d;
    g.n1 = g.k * g.nc + 1;
    g.N  = (long long)g.n1 * g.n1 * (g.dim == 3 ? g.n1 : 1);
    g.h  = 1.0 / g.nc;
    return g;
  }

  // operator coefficients handed to kernels by value
  struct OpDev
  {
    int    nb;
    double cm[SPIRK_MAX_BLOCKS];                      // mass[b] * h^dim  (REAL)
    double cl[SPIRK_MAX_BLOCKS];                      // laplace[b] * h^(dim-2)
    double cc[SPIRK_MAX_BLOCKS * SPIRK_MAX_BLOCKS];   // coupling * h^dim (COUPLED)
  };

  thread_local extern std::string g_last_error;
  int set_error(int code, const std::string &msg);

#define SPIRK_CUDA(call)                                                                        \
  do                                                                                            \
    {                                                                                           \
      cudaError_t err__ = (call);                                                               \
      if (err__ != cudaSuccess)                                                                 \
        return spirk::set_error(SPIRK_ERR_DEVICE,                                               \
                                std::string(#call) + ": " + cudaGetErrorString(err__));         \
    }                                                                                           \
  while (0)

#define SPIRK_LAUNCH_CHECK(ctx)                                                                 \
  do                                                                                            \
    {                                                                                           \
      (ctx)->launches++;                                                                        \
      cudaError_t err__ = cudaGetLastError();                                                   \
      if (err__ != cudaSuccess)                                                                 \
        return spirk::set_error(SPIRK_ERR_DEVICE, std::string("kernel launch: ") +              \
                                                    cudaGetErrorString(err__));                 \
    }                                                                                           \
  while (0)
} // namespace spirk

struct spirk_ctx
{
  int          device   = 0;
  cudaStream_t stream   = nullptr;
  long long    launches = 0;
  int          n_sms    = 148;
  // reduction scratch: per-block partials + result slots (device), pinned host mirror
  double *d_partials = nullptr; // capacity n_partials
  int     n_partials = 0;
  double *d_result   = nullptr; // 64 doubles
  double *h_result   = nullptr; // pinned, 64 doubles
  // operator scratch (A x for unfused epilogues)
  double *d_scratch   = nullptr;
  size_t  scratch_cap = 0;
  // small 1-D table scratch
  double *d_tab   = nullptr;
  size_t  tab_cap = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  spirk_comm *reduction_comm = nullptr;
  int         opt_apply_variant = 0;
};

namespace spirk
{
  int ensure_scratch(spirk_ctx *ctx, size_t n);
  int ensure_tab(spirk_ctx *ctx, size_t n);
  int upload_fe_constants();
  // host copy of the reference matrices as uploaded (exactly persymmetric), (k+1)^2 entries each
  void fe_host_sym(int k, double *Mh, double *Kh);
  // reduce-to-host helper: sums ctx->d_partials[0..count) per result slot
  int finish_reduction(spirk_ctx *ctx, int n_results, int n_blocks, double *host_out);
} // namespace spirk
