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

  // geometry handed to kernels by value.  A level may be a z-SLAB of the mesh (3-D; spirk_level::slab): cell layers
  // [L_lo, L_hi) of the nc, owned node planes [zo0, zo1) (a vertex plane between two slabs belongs to the upper one, the top
  // plane of the domain to the last slab).  Vectors of a slab level hold their owned planes contiguously; pointers handed
  // to the library point at the first OWNED entry, and SPIRK_SLAB_PAD_LO planes below / SPIRK_SLAB_PAD_HI planes above it
  // are ghost planes (filled by spirk_halo_exchange).  N = owned entries.  Unpartitioned: L = [0, nc), zo = [0, n1).
  struct Geo
  {
    int       dim, k, n, nc, n1;
    long long N;
    double    h;
    int       col_rank, col_size, coarse_replicated;
    int       L_lo, L_hi, zo0, zo1, gh_lo, gh_hi;
    long long plane;
  };

  inline Geo make_geo(const spirk_level *l)
  {
    Geo g;
    g.dim = l->dim, g.k = l->degree, g.n = l->degree + 1, g.nc = l->n_cells_1d;
    g.n1    = g.k * g.nc + 1;
    g.plane = (g.dim == 3) ? (long long)g.n1 * g.n1 : g.n1;
    g.h     = 1.0 / g.nc;
    g.col_size = (l->slab >> 8) & 0xff, g.col_rank = l->slab & 0xff, g.coarse_replicated = (l->slab >> 16) & 1;
    if (g.col_size <= 1)
      g.col_size = 1, g.col_rank = 0;
    g.L_lo = (int)((long long)g.nc * g.col_rank / g.col_size), g.L_hi = (int)((long long)g.nc * (g.col_rank + 1) / g.col_size);
    g.zo0 = g.k * g.L_lo, g.zo1 = (g.L_hi == g.nc) ? g.n1 : g.k * g.L_hi;
    g.gh_lo = (g.col_size > 1) ? SPIRK_SLAB_PAD_LO(g.k) : 0, g.gh_hi = (g.col_size > 1) ? SPIRK_SLAB_PAD_HI : 0;
    g.N     = (g.dim == 3) ? g.plane * (g.zo1 - g.zo0) : (long long)g.n1 * g.n1;
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
  // work queue of the plane-streaming cell operator (op_v3.cuh): [0] next item, [1] finished CTAs, [4...] boundary states;
  // partial sums exchanged at the range boundaries
  int    *d_v3_sched     = nullptr;
  size_t  v3_sched_cap   = 0;
  double *d_v3_carry     = nullptr;
  size_t  v3_carry_cap   = 0;
  std::vector<void *> retired; // outgrown queue arrays, kept alive for captured graphs
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  spirk_comm *reduction_comm = nullptr;
  // spirk_ctx_set_option knobs (documented in include/spirk_b200.h); defaults = the validated fastest configuration
  int opt_apply_variant  = 0;  // "apply_variant": 0 plane streaming (op_v3), 2 / 3 tile columns (op_v2), 1 general cell kernel only
  int opt_v3_schedule    = -1; // "v3_schedule": -1 / 2 work queue + carry exchange (default), 0 even static split, 1 z-lockstep,
                               // 3 the round-1 heuristic (lockstep for vectors beyond the L2 capacity, else even split)
  int opt_v3_chunk       = 0;  // "v3_chunk": layers per work item (0 = 8)
  int opt_v3_tail        = 1;  // "v3_tail": 1 = long ranges first, short ranges for the tail of a launch (0 = equal ranges)
  int opt_transfer_variant = 0; // "transfer_variant": 0 owner-computes 1-D sweeps (restriction z, y, x), 1 cell-based kernels (restriction
                                // with atomics), 2 owner-computes sweeps with the restriction in the order x, y, z
  int opt_v3_grid        = 0;  // "v3_grid": 0 = all co-resident CTAs, > 0 = this many CTAs (even split)
  int opt_v3_smem_pad_kb = 0;  // "v3_smem_pad_kb": extra dynamic shared memory per CTA (limits the CTAs per SM; experiments)
  int opt_v3_npt         = 0;  // "v3_npt": nodes per y+z thread on 8 x 8 tiles, 0 = per mode (4 apply, 2 fused epilogues)
  int opt_v3_small_below = 32; // "v3_small_below": levels with fewer cells per direction use 4 x 4-cell tiles
  int opt_v3_l2promo     = 0;  // "v3_l2promo": CUtensorMapL2promotion of the staging maps (0 none .. 3 256 B)
};

namespace spirk
{
  int ensure_scratch(spirk_ctx *ctx, size_t n);
  int ensure_tab(spirk_ctx *ctx, size_t n);
  int ensure_v3_queue(spirk_ctx *ctx, size_t n_states, size_t n_carry);
  int upload_fe_constants();
  // host copy of the reference matrices as uploaded (exactly persymmetric), (k+1)^2 entries each
  void fe_host_sym(int k, double *Mh, double *Kh);
  // reduce-to-host helper: sums ctx->d_partials[0..count) per result slot
  int finish_reduction(spirk_ctx *ctx, int n_results, int n_blocks, double *host_out);
} // namespace spirk
