// Implementation of include/spirk_b200.h with hand-written CUDA kernels for sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC ...
// There is NO CPU fallback in this library: every entry point needs a CUDA device.

#include <cmath>
#include <cstring>
#include <mutex>

#include "fe1d.h"
#include "misc_kernels.cuh"
#include "nccl_dl.h"
#include "op_v3.cuh"

namespace spirk
{
  thread_local std::string g_last_error;
  int                      set_error(int code, const std::string &msg)
  {
    g_last_error = msg;
    return code;
  }

  static Fe1D       g_fe[SPIRK_MAX_DEGREE + 1];
  static std::mutex g_fe_mutex;
  static bool       g_fe_host_ready = false;

  static void init_fe_host()
  {
    std::lock_guard<std::mutex> lock(g_fe_mutex);
    if (!g_fe_host_ready)
      {
        for (int k = 1; k <= SPIRK_MAX_DEGREE; ++k)
          g_fe[k] = make_fe1d(k);
        g_fe_host_ready = true;
      }
  }

  int upload_fe_constants()
  {
    init_fe_host();
    static FeConst all[SPIRK_MAX_DEGREE + 1];
    std::memset(all, 0, sizeof(all));
    for (int k = 1; k <= SPIRK_MAX_DEGREE; ++k)
      {
        const Fe1D &f = g_fe[k];
        const int   n = f.n;
        // the reference matrices are persymmetric (the GLL nodes are symmetric about the cell centre); make
        // that exact in the device copy so kernels may share the value of mirrored entries (op_v3.cuh)
        for (int i = 0; i < n; ++i)
          for (int j = 0; j < n; ++j)
            {
              all[k].Mh[i * n + j] = 0.5 * (f.Mh[i * n + j] + f.Mh[(k - i) * n + (k - j)]);
              all[k].Kh[i * n + j] = 0.5 * (f.Kh[i * n + j] + f.Kh[(k - i) * n + (k - j)]);
            }
        for (int i = 0; i < (2 * k + 1) * n; ++i)
          all[k].P[i] = f.P[i];
        for (int i = 0; i < (k + 2) * n; ++i)
          all[k].Be[i] = f.Be[i];
        for (int i = 0; i < k + 2; ++i)
          all[k].xe[i] = f.xe[i], all[k].we[i] = f.we[i];
        for (int i = 0; i < n; ++i)
          all[k].nodes[i] = f.nodes[i];
        all[k].Mv = all[k].Mh[k * n + k] + all[k].Mh[0], all[k].Kv = all[k].Kh[k * n + k] + all[k].Kh[0];
      }
    SPIRK_CUDA(cudaMemcpyToSymbol(c_fe, all, sizeof(all)));
    // the copies of the separately compiled plane-streaming kernels
    if (int e = v3_upload_constants_mode0(all))
      return e;
    if (int e = v3_upload_constants_mode1(all))
      return e;
    if (int e = v3_upload_constants_mode2(all))
      return e;
    if (int e = v3_upload_constants_mode3(all))
      return e;
    if (int e = v3_upload_constants_mode4(all))
      return e;
    if (int e = v3_upload_constants_mode5(all))
      return e;
    if (int e = v3_upload_constants_mode6(all))
      return e;
    return SPIRK_OK;
  }

  void fe_host_sym(int k, double *Mh, double *Kh)
  {
    init_fe_host();
    const Fe1D &f = g_fe[k];
    const int   n = f.n;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j)
        {
          Mh[i * n + j] = 0.5 * (f.Mh[i * n + j] + f.Mh[(k - i) * n + (k - j)]);
          Kh[i * n + j] = 0.5 * (f.Kh[i * n + j] + f.Kh[(k - i) * n + (k - j)]);
        }
  }

  int ensure_scratch(spirk_ctx *ctx, size_t n)
  {
    if (ctx->scratch_cap < n)
      {
        if (ctx->d_scratch)
          ctx->retired.push_back(ctx->d_scratch); // a captured graph may still hold the old pointer
        ctx->d_scratch = nullptr, ctx->scratch_cap = 0;
        SPIRK_CUDA(cudaMalloc(&ctx->d_scratch, n * sizeof(double)));
        ctx->scratch_cap = n;
      }
    return SPIRK_OK;
  }

  int ensure_tab(spirk_ctx *ctx, size_t n)
  {
    if (ctx->tab_cap < n)
      {
        if (ctx->d_tab)
          {
            SPIRK_CUDA(cudaStreamSynchronize(ctx->stream));
            SPIRK_CUDA(cudaFree(ctx->d_tab));
          }
        ctx->d_tab = nullptr, ctx->tab_cap = 0;
        SPIRK_CUDA(cudaMalloc(&ctx->d_tab, n * sizeof(double)));
        ctx->tab_cap = n;
      }
    return SPIRK_OK;
  }

  int ensure_v3_queue(spirk_ctx *ctx, size_t n_states, size_t n_carry)
  {
    // Grown arrays never replace the old ones in place: a captured CUDA graph may hold the old pointers, so those stay
    // valid (and zeroed, as every launch leaves them) until the context is destroyed.
    if (ctx->v3_sched_cap < n_states + 4)
      {
        if (ctx->d_v3_sched)
          ctx->retired.push_back(ctx->d_v3_sched);
        ctx->d_v3_sched = nullptr, ctx->v3_sched_cap = 0;
        const size_t cap = 2 * (n_states + 4);
        SPIRK_CUDA(cudaMalloc(&ctx->d_v3_sched, cap * sizeof(int)));
        SPIRK_CUDA(cudaMemsetAsync(ctx->d_v3_sched, 0, cap * sizeof(int), ctx->stream)); // the kernels leave it zeroed
        ctx->v3_sched_cap = cap;
      }
    if (ctx->v3_carry_cap < n_carry)
      {
        if (ctx->d_v3_carry)
          ctx->retired.push_back(ctx->d_v3_carry);
        ctx->d_v3_carry = nullptr, ctx->v3_carry_cap = 0;
        const size_t cap = n_carry + n_carry / 2;
        SPIRK_CUDA(cudaMalloc(&ctx->d_v3_carry, cap * sizeof(double)));
        ctx->v3_carry_cap = cap;
      }
    return SPIRK_OK;
  }

  static int check_level(const spirk_level *l)
  {
    if (!l || (l->dim != 2 && l->dim != 3) || l->n_cells_1d < 1)
      return set_error(SPIRK_ERR_INVALID, "bad level descriptor");
    if (l->degree < 1 || l->degree > SPIRK_MAX_DEGREE)
      return set_error(SPIRK_ERR_UNSUPPORTED, "degree must be 1..6");
    const int col_size = (l->slab >> 8) & 0xff, col_rank = l->slab & 0xff;
    if (col_size > 1 && (l->dim != 3 || l->n_cells_1d % col_size != 0 || col_rank >= col_size))
      return set_error(SPIRK_ERR_INVALID, "z-slab levels: 3-D, n_cells_1d a multiple of the number of slabs, rank < size");
    return SPIRK_OK;
  }

  static int check_op(const spirk_opdesc *op)
  {
    if (!op || op->nb < 1 || op->nb > SPIRK_MAX_BLOCKS || (op->kind != SPIRK_OP_REAL && op->kind != SPIRK_OP_COUPLED))
      return set_error(SPIRK_ERR_INVALID, "bad operator descriptor");
    return SPIRK_OK;
  }

  OpDev make_opdev(const Geo &g, const spirk_opdesc *op)
  {
    OpDev        d;
    const double hd = std::pow(g.h, g.dim), hl = std::pow(g.h, g.dim - 2);
    d.nb = op->nb;
    for (int b = 0; b < op->nb; ++b)
      {
        d.cm[b] = op->mass[b] * hd;
        d.cl[b] = op->laplace[b] * hl;
        for (int j = 0; j < op->nb; ++j)
          d.cc[b * op->nb + j] = (op->kind == SPIRK_OP_COUPLED) ? op->coupling[b * op->nb + j] * hd : 0.0;
      }
    return d;
  }

  static inline int grid_for(const spirk_ctx *ctx, long long n, int threads, int per_sm = 8)
  {
    long long blocks = (n + threads - 1) / threads;
    long long cap    = (long long)ctx->n_sms * per_sm;
    return (int)std::max<long long>(1, std::min(blocks, cap));
  }

  template <int K>
  static int launch_apply_v1(spirk_ctx *ctx, const Geo &g, const OpDev &od, bool coupled, double *dst, const double *src,
                             long long stride)
  {
    using C = CfgV1<K>;
    k_init_dst<<<grid_for(ctx, g.N * od.nb, 256), 256, 0, ctx->stream>>>(g, od.nb, dst, src, stride);
    SPIRK_LAUNCH_CHECK(ctx);
    if (g.dim == 3)
      {
        const long long ncells = (long long)g.nc * g.nc * g.nc, nbatch = (ncells + C::CPB3 - 1) / C::CPB3;
        const int       grid   = (int)std::min<long long>(nbatch, (long long)ctx->n_sms * 8);
        if (coupled)
          k_apply3d_v1<K, true><<<grid, C::T3, 0, ctx->stream>>>(g, od, dst, src, stride);
        else
          k_apply3d_v1<K, false><<<grid, C::T3, 0, ctx->stream>>>(g, od, dst, src, stride);
      }
    else
      {
        const long long ncells = (long long)g.nc * g.nc, nbatch = (ncells + C::CPB2 - 1) / C::CPB2;
        const int       grid   = (int)std::min<long long>(nbatch, (long long)ctx->n_sms * 8);
        if (coupled)
          k_apply2d_v1<K, true><<<grid, C::T2, 0, ctx->stream>>>(g, od, dst, src, stride);
        else
          k_apply2d_v1<K, false><<<grid, C::T2, 0, ctx->stream>>>(g, od, dst, src, stride);
      }
    SPIRK_LAUNCH_CHECK(ctx);
    return SPIRK_OK;
  }

#define SPIRK_DISPATCH_K(k, expr)                                           \
  switch (k)                                                                \
    {                                                                       \
      case 1: { constexpr int K = 1; expr; } break;                         \
      case 2: { constexpr int K = 2; expr; } break;                         \
      case 3: { constexpr int K = 3; expr; } break;                         \
      case 4: { constexpr int K = 4; expr; } break;                         \
      case 5: { constexpr int K = 5; expr; } break;                         \
      case 6: { constexpr int K = 6; expr; } break;                         \
      default: return set_error(SPIRK_ERR_UNSUPPORTED, "degree");          \
    }

  // dst = A src with the variant selected by the context; mode-specific fused paths in op_v2
  static int apply_any(spirk_ctx *ctx, const Geo &g, const spirk_opdesc *op, double *dst, const double *src, long long stride)
  {
    if (g.col_size > 1)
      return set_error(SPIRK_ERR_UNSUPPORTED, "z-slab levels need the plane-streaming cell operator (3-D, Q4, >= 8 cells per direction)");
    const OpDev od      = make_opdev(g, op);
    const bool  coupled = op->kind == SPIRK_OP_COUPLED;
    int         st      = SPIRK_OK;
    SPIRK_DISPATCH_K(g.k, st = launch_apply_v1<K>(ctx, g, od, coupled, dst, src, stride));
    return st;
  }

  int finish_reduction(spirk_ctx *ctx, int n_results, int n_blocks, double *host_out)
  {
    for (int r = 0; r < n_results; ++r)
      {
        k_finish<<<1, RT, 0, ctx->stream>>>(ctx->d_partials + (size_t)r * n_blocks, n_blocks, ctx->d_result + r);
        SPIRK_LAUNCH_CHECK(ctx);
      }
    if (ctx->reduction_comm)
      if (int e = spirk_comm_allreduce_sum(ctx, ctx->reduction_comm, ctx->d_result, n_results))
        return e;
    SPIRK_CUDA(cudaMemcpyAsync(ctx->h_result, ctx->d_result, n_results * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SPIRK_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int r = 0; r < n_results; ++r)
      host_out[r] = ctx->h_result[r];
    return SPIRK_OK;
  }
} // namespace spirk

using namespace spirk;

struct spirk_comm
{
  ncclComm_t comm = nullptr;
  int        rank = 0, n_ranks = 1;
};

#define SPIRK_NCCL(call)                                                                                     \
  do                                                                                                         \
    {                                                                                                        \
      const NcclApi &nccl = nccl_api();                                                                      \
      if (!nccl.ok)                                                                                          \
        return set_error(SPIRK_ERR_COMM, nccl.error);                                                        \
      ncclResult_t r__ = (call);                                                                             \
      if (r__ != ncclSuccess)                                                                                \
        return set_error(SPIRK_ERR_COMM, std::string(#call) + ": " + nccl.GetErrorString(r__));             \
    }                                                                                                        \
  while (0)

extern "C" {

const char *spirk_backend(void) { return "cuda-sm_100a"; }
const char *spirk_last_error(void) { return g_last_error.c_str(); }

int spirk_ctx_create(spirk_ctx **out, int device)
{
  if (!out)
    return set_error(SPIRK_ERR_INVALID, "null ctx pointer");
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
    return set_error(SPIRK_ERR_DEVICE, "no CUDA device available — this library has no CPU fallback");
  if (device < 0 || device >= count)
    return set_error(SPIRK_ERR_INVALID, "device index out of range");
  SPIRK_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  SPIRK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return set_error(SPIRK_ERR_DEVICE, "kernels are compiled for sm_100a only");
  spirk_ctx *ctx = new spirk_ctx();
  ctx->device    = device;
  ctx->n_sms     = prop.multiProcessorCount;
  SPIRK_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  ctx->n_partials = ctx->n_sms * 8 * 4;
  SPIRK_CUDA(cudaMalloc(&ctx->d_partials, ctx->n_partials * sizeof(double)));
  SPIRK_CUDA(cudaMalloc(&ctx->d_result, 64 * sizeof(double)));
  SPIRK_CUDA(cudaMallocHost(&ctx->h_result, 64 * sizeof(double)));
  SPIRK_CUDA(cudaEventCreate(&ctx->ev0));
  SPIRK_CUDA(cudaEventCreate(&ctx->ev1));
  if (int e = upload_fe_constants())
    return e;
  *out = ctx;
  return SPIRK_OK;
}

int spirk_ctx_destroy(spirk_ctx *ctx)
{
  if (!ctx)
    return SPIRK_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(ctx->d_partials);
  cudaFree(ctx->d_result);
  cudaFreeHost(ctx->h_result);
  cudaFree(ctx->d_scratch);
  cudaFree(ctx->d_tab);
  cudaFree(ctx->d_v3_sched);
  cudaFree(ctx->d_v3_carry);
  for (void *p : ctx->retired)
    cudaFree(p);
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  return SPIRK_OK;
}

int spirk_ctx_sync(spirk_ctx *ctx)
{
  SPIRK_CUDA(cudaStreamSynchronize(ctx->stream));
  return SPIRK_OK;
}
long long spirk_ctx_launch_count(spirk_ctx *ctx) { return ctx->launches; }
int       spirk_ctx_timer_begin(spirk_ctx *ctx)
{
  SPIRK_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  return SPIRK_OK;
}
int spirk_ctx_timer_end(spirk_ctx *ctx, double *ms)
{
  SPIRK_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  SPIRK_CUDA(cudaEventSynchronize(ctx->ev1));
  float f = 0;
  SPIRK_CUDA(cudaEventElapsedTime(&f, ctx->ev0, ctx->ev1));
  *ms = f;
  return SPIRK_OK;
}
int spirk_ctx_set_option(spirk_ctx *ctx, const char *name, int value)
{
  struct
  {
    const char *name;
    int        *slot;
  } const table[] = {{"apply_variant", &ctx->opt_apply_variant},   {"v3_schedule", &ctx->opt_v3_schedule},
                     {"v3_grid", &ctx->opt_v3_grid},               {"v3_smem_pad_kb", &ctx->opt_v3_smem_pad_kb},
                     {"v3_npt", &ctx->opt_v3_npt},                 {"v3_small_below", &ctx->opt_v3_small_below},
                     {"v3_l2promo", &ctx->opt_v3_l2promo},         {"v3_chunk", &ctx->opt_v3_chunk},
                     {"v3_tail", &ctx->opt_v3_tail},
                     {"transfer_variant", &ctx->opt_transfer_variant}};
  for (const auto &t : table)
    if (std::strcmp(name, t.name) == 0)
      {
        *t.slot = value;
        return SPIRK_OK;
      }
  return set_error(SPIRK_ERR_INVALID, std::string("unknown option ") + name);
}

// ------------------------------------------------------------------------------- graphs
struct spirk_graph
{
  cudaGraph_t     graph = nullptr;
  cudaGraphExec_t exec  = nullptr;
  size_t          n_kernels = 0; // kernel nodes of the captured sequence (for the launch count)
};
int spirk_graph_begin(spirk_ctx *ctx)
{
  SPIRK_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
  return SPIRK_OK;
}
int spirk_graph_end(spirk_ctx *ctx, spirk_graph **out)
{
  spirk_graph *g = new spirk_graph();
  cudaError_t  e = cudaStreamEndCapture(ctx->stream, &g->graph);
  if (e == cudaSuccess)
    e = cudaGraphInstantiate(&g->exec, g->graph, 0);
  if (e != cudaSuccess)
    {
      if (g->graph)
        cudaGraphDestroy(g->graph);
      delete g;
      cudaGetLastError();
      return set_error(SPIRK_ERR_DEVICE, std::string("graph capture: ") + cudaGetErrorString(e));
    }
  {
    size_t n = 0;
    if (cudaGraphGetNodes(g->graph, nullptr, &n) == cudaSuccess && n > 0)
      {
        std::vector<cudaGraphNode_t> nodes(n);
        if (cudaGraphGetNodes(g->graph, nodes.data(), &n) == cudaSuccess)
          for (size_t i = 0; i < n; ++i)
            {
              cudaGraphNodeType t;
              if (cudaGraphNodeGetType(nodes[i], &t) == cudaSuccess && t == cudaGraphNodeTypeKernel)
                g->n_kernels++;
            }
      }
  }
  *out = g;
  return SPIRK_OK;
}
int spirk_graph_launch(spirk_ctx *ctx, spirk_graph *g)
{
  SPIRK_CUDA(cudaGraphLaunch(g->exec, ctx->stream));
  ctx->launches += (long long)std::max<size_t>(1, g->n_kernels); // every replayed kernel counts
  return SPIRK_OK;
}
int spirk_graph_destroy(spirk_graph *g)
{
  if (g)
    {
      if (g->exec)
        cudaGraphExecDestroy(g->exec);
      if (g->graph)
        cudaGraphDestroy(g->graph);
      delete g;
    }
  return SPIRK_OK;
}

int spirk_malloc(spirk_ctx *ctx, double **ptr, size_t n)
{
  (void)ctx;
  if (cudaMalloc(ptr, (n ? n : 1) * sizeof(double)) != cudaSuccess)
    {
      cudaGetLastError();
      return set_error(SPIRK_ERR_NOMEM, "cudaMalloc failed");
    }
  SPIRK_CUDA(cudaMemsetAsync(*ptr, 0, (n ? n : 1) * sizeof(double), ctx->stream));
  return SPIRK_OK;
}
int spirk_free(spirk_ctx *ctx, double *ptr)
{
  SPIRK_CUDA(cudaStreamSynchronize(ctx->stream));
  SPIRK_CUDA(cudaFree(ptr));
  return SPIRK_OK;
}
int spirk_copy_h2d(spirk_ctx *ctx, double *dst, const double *src, size_t n)
{
  SPIRK_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  SPIRK_CUDA(cudaStreamSynchronize(ctx->stream));
  return SPIRK_OK;
}
int spirk_copy_d2h(spirk_ctx *ctx, double *dst, const double *src, size_t n)
{
  SPIRK_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  SPIRK_CUDA(cudaStreamSynchronize(ctx->stream));
  return SPIRK_OK;
}
int spirk_malloc_host(spirk_ctx *, double **ptr, size_t n)
{
  SPIRK_CUDA(cudaMallocHost(ptr, (n ? n : 1) * sizeof(double)));
  return SPIRK_OK;
}
int spirk_free_host(spirk_ctx *, double *ptr)
{
  SPIRK_CUDA(cudaFreeHost(ptr));
  return SPIRK_OK;
}

long long spirk_level_n_dofs(const spirk_level *lvl) { return make_geo(lvl).N; }

// ------------------------------------------------------------------------------- operator
// fast-path dispatch: variant 0 (default) = plane-streaming kernel (op_v3.cuh), 2 / 3 = pipelined
// tile-column kernel (op_v2.cuh), 1 = general cell kernel only (op_v1.cuh)
static int fast_apply(spirk_ctx *ctx, const Geo &g, const spirk_opdesc *op, V2Mode mode, double *dst, const double *src,
                      const double *x_old, const double *rhs, const double *dinv, long long stride, const double *f1,
                      const double *f2, const double *diag_mass = nullptr, const double *diag_laplace = nullptr)
{
  if (ctx->opt_apply_variant == 1)
    return SPIRK_ERR_UNSUPPORTED;
  if (ctx->opt_apply_variant == 0)
    {
      int st = v3_apply(ctx, g, op, mode, dst, src, x_old, rhs, dinv, stride, f1, f2, nullptr, nullptr, diag_mass, diag_laplace);
      if (st != SPIRK_ERR_UNSUPPORTED)
        return st;
    }
  if (diag_mass || diag_laplace)
    return SPIRK_ERR_UNSUPPORTED; // (the tile-column kernel only knows the operator's own diagonal)
  return v2_apply(ctx, g, op, mode, dst, src, x_old, rhs, dinv, stride, f1, f2);
}
// dst = A src into a scratch vector whose blocks sit at stride N: the fastest kernel that covers the operator
static int apply_fast_or_any(spirk_ctx *ctx, const Geo &g, const spirk_opdesc *op, double *dst, const double *src, long long src_stride,
                             long long dst_stride)
{
  if (src_stride == dst_stride)
    {
      int st = fast_apply(ctx, g, op, V2_APPLY, dst, src, nullptr, nullptr, nullptr, src_stride, nullptr, nullptr);
      if (st != SPIRK_ERR_UNSUPPORTED)
        return st;
    }
  return apply_any(ctx, g, op, dst, src, dst_stride);
}

int spirk_op_apply(spirk_ctx *ctx, const spirk_level *lvl, const spirk_opdesc *op, double *dst, const double *src,
                   long long stride)
{
  if (int e = check_level(lvl))
    return e;
  if (int e = check_op(op))
    return e;
  if (dst == src)
    return set_error(SPIRK_ERR_INVALID, "op_apply: dst must not alias src");
  const Geo g = make_geo(lvl);
  {
    int st = fast_apply(ctx, g, op, V2_APPLY, dst, src, nullptr, nullptr, nullptr, stride, nullptr, nullptr);
    if (st != SPIRK_ERR_UNSUPPORTED)
      return st;
  }
  return apply_any(ctx, g, op, dst, src, stride);
}

int spirk_op_apply_km(spirk_ctx *ctx, const spirk_level *lvl, int nb, double *dst, const double *v, const double *w, long long stride,
                      const double *laplace, const double *mass)
{
  if (int e = check_level(lvl))
    return e;
  if (nb < 1 || nb > SPIRK_MAX_BLOCKS || dst == v || dst == w)
    return set_error(SPIRK_ERR_INVALID, "op_apply_km: block count / aliasing");
  const Geo g = make_geo(lvl);
  if (ctx->opt_apply_variant == 0)
    {
      int st = v3_apply_km(ctx, g, nb, dst, v, w, stride, laplace, mass);
      if (st != SPIRK_ERR_UNSUPPORTED)
        return st;
    }
  // general path: dst = laplace K v, scratch = mass M w, dst += scratch (Dirichlet rows: dst = v)
  spirk_opdesc dk, dm;
  std::memset(&dk, 0, sizeof(dk)), std::memset(&dm, 0, sizeof(dm));
  dk.kind = dm.kind = SPIRK_OP_REAL, dk.nb = dm.nb = nb;
  for (int b = 0; b < nb; ++b)
    dk.laplace[b] = laplace[b], dm.mass[b] = mass[b];
  if (int e = spirk_op_apply(ctx, lvl, &dk, dst, v, stride))
    return e;
  if (int e = ensure_scratch(ctx, (size_t)stride * nb)) // (one stride for source and destination blocks)
    return e;
  if (int e = spirk_op_apply(ctx, lvl, &dm, ctx->d_scratch, w, stride))
    return e;
  k_add_interior<<<grid_for(ctx, g.N * nb, 256), 256, 0, ctx->stream>>>(g, nb, dst, stride, ctx->d_scratch, stride);
  SPIRK_LAUNCH_CHECK(ctx);
  return SPIRK_OK;
}

int spirk_op_residual(spirk_ctx *ctx, const spirk_level *lvl, const spirk_opdesc *op, double *dst, const double *rhs,
                      const double *src, long long stride)
{
  if (int e = check_level(lvl))
    return e;
  if (int e = check_op(op))
    return e;
  if (dst == src)
    return set_error(SPIRK_ERR_INVALID, "op_residual: dst must not alias src");
  const Geo g = make_geo(lvl);
  {
    int st = fast_apply(ctx, g, op, V2_RESIDUAL, dst, src, nullptr, rhs, nullptr, stride, nullptr, nullptr);
    if (st != SPIRK_ERR_UNSUPPORTED)
      return st;
  }
  if (int e = ensure_scratch(ctx, (size_t)g.N * op->nb))
    return e;
  // unfused: A src through the fastest kernel that covers the operator (coupled pairs: the plane-streaming apply), then
  // the pointwise epilogue
  if (int e = apply_fast_or_any(ctx, g, op, ctx->d_scratch, src, stride, g.N))
    return e;
  k_residual_epilogue<<<grid_for(ctx, g.N * op->nb, 256), 256, 0, ctx->stream>>>(g.N, op->nb, dst, rhs, ctx->d_scratch, stride, g.N);
  SPIRK_LAUNCH_CHECK(ctx);
  return SPIRK_OK;
}

// mass factor of block b's own (diagonal) term
static double own_mass(const spirk_opdesc *op, int b) { return op->kind == SPIRK_OP_REAL ? op->mass[b] : op->coupling[b * op->nb + b]; }

// one Chebyshev iteration; dinv == NULL: the inverse diagonal of diag_mass[b] M + diag_laplace[b] K (NULL: of the operator)
static int cheb_step_impl(spirk_ctx *ctx, const spirk_level *lvl, const spirk_opdesc *op, double *x_new, const double *x,
                          const double *x_old, const double *rhs, const double *dinv, const double *diag_mass,
                          const double *diag_laplace, long long stride, const double *f1, const double *f2)
{
  if (int e = check_level(lvl))
    return e;
  if (int e = check_op(op))
    return e;
  if (x_new == x)
    return set_error(SPIRK_ERR_INVALID, "op_cheb_step: x_new must not alias x (it may alias x_old)");
  const Geo g = make_geo(lvl);
  {
    int st = fast_apply(ctx, g, op, V2_CHEB, x_new, x, x_old, rhs, dinv, stride, f1, f2, dinv ? nullptr : diag_mass,
                        dinv ? nullptr : diag_laplace);
    if (st != SPIRK_ERR_UNSUPPORTED)
      return st;
  }
  // general path: A x into scratch, then the pointwise update; dinv == NULL -> the operator's own
  // inverse diagonal (COUPLED: of block b's own term coupling[b][b] M + laplace[b] K), materialised behind A x in the
  // scratch buffer
  const bool own_dinv = (dinv == nullptr);
  if (int e = ensure_scratch(ctx, (size_t)g.N * op->nb * (own_dinv ? 2 : 1)))
    return e;
  if (own_dinv)
    {
      double *d = ctx->d_scratch + (size_t)g.N * op->nb;
      for (int b = 0; b < op->nb; ++b)
        if (int e = spirk_op_inverse_diagonal(ctx, lvl, d + (size_t)b * g.N, diag_mass ? diag_mass[b] : own_mass(op, b),
                                              diag_laplace ? diag_laplace[b] : op->laplace[b]))
          return e;
    }
  if (int e = apply_fast_or_any(ctx, g, op, ctx->d_scratch, x, stride, g.N))
    return e;
  ChebFactors f;
  for (int b = 0; b < op->nb; ++b)
    f.f1[b] = f1[b], f.f2[b] = f2[b];
  k_cheb_epilogue<<<grid_for(ctx, g.N * op->nb, 256), 256, 0, ctx->stream>>>(
    g.N, op->nb, x_new, x, x_old, rhs, own_dinv ? ctx->d_scratch + (size_t)g.N * op->nb : dinv, ctx->d_scratch, stride, g.N, f,
    own_dinv ? g.N : stride);
  SPIRK_LAUNCH_CHECK(ctx);
  return SPIRK_OK;
}

int spirk_op_cheb_step(spirk_ctx *ctx, const spirk_level *lvl, const spirk_opdesc *op, double *x_new, const double *x,
                       const double *x_old, const double *rhs, const double *dinv, long long stride, const double *f1,
                       const double *f2)
{
  return cheb_step_impl(ctx, lvl, op, x_new, x, x_old, rhs, dinv, nullptr, nullptr, stride, f1, f2);
}

int spirk_op_cheb_step_diag(spirk_ctx *ctx, const spirk_level *lvl, const spirk_opdesc *op, double *x_new, const double *x,
                            const double *x_old, const double *rhs, const double *diag_mass, const double *diag_laplace,
                            long long stride, const double *f1, const double *f2)
{
  if (!diag_mass || !diag_laplace)
    return set_error(SPIRK_ERR_INVALID, "op_cheb_step_diag: the coefficients of the diagonal are required");
  return cheb_step_impl(ctx, lvl, op, x_new, x, x_old, rhs, nullptr, diag_mass, diag_laplace, stride, f1, f2);
}

int spirk_op_fuses_own_diagonal(spirk_ctx *ctx, const spirk_level *lvl, const spirk_opdesc *op)
{
  if (!lvl || !op || ctx->opt_apply_variant != 0 || (op->kind != SPIRK_OP_REAL && !(op->kind == SPIRK_OP_COUPLED && op->nb == 2)))
    return 0;
  const Geo g = make_geo(lvl);
  return (g.dim == 3 && g.k == 4 && g.nc % 4 == 0 && g.nc >= 8) ? 1 : 0; // the shapes v3_apply covers
}

static int cheb_first_impl(spirk_ctx *ctx, const spirk_level *lvl, const spirk_opdesc *op, double *x1, double *x2, const double *rhs,
                           const double *diag_mass, const double *diag_laplace, long long stride, const double *f0, const double *f1,
                           const double *f2)
{
  if (int e = check_level(lvl))
    return e;
  if (int e = check_op(op))
    return e;
  if (x1 == x2 || x1 == rhs || x2 == rhs)
    return set_error(SPIRK_ERR_INVALID, "op_cheb_first: x1, x2 and rhs must be distinct");
  const Geo g = make_geo(lvl);
  if (ctx->opt_apply_variant == 0)
    {
      int st = v3_apply(ctx, g, op, V2_CHEB_FIRST, x2, rhs, nullptr, nullptr, nullptr, stride, f1, f2, f0, x1, diag_mass, diag_laplace);
      if (st != SPIRK_ERR_UNSUPPORTED)
        return st;
    }
  // general path: D^-1 into x2, x1 = f0 D^-1 rhs, then the ordinary Chebyshev step with x_old = 0
  for (int b = 0; b < op->nb; ++b)
    if (int e = spirk_op_inverse_diagonal(ctx, lvl, x2 + (size_t)b * stride, diag_mass ? diag_mass[b] : own_mass(op, b),
                                          diag_laplace ? diag_laplace[b] : op->laplace[b]))
      return e;
  if (int e = spirk_vec_scale_pointwise(ctx, op->nb, g.N, x1, x2, rhs, stride, f0))
    return e;
  return cheb_step_impl(ctx, lvl, op, x2, x1, nullptr, rhs, nullptr, diag_mass, diag_laplace, stride, f1, f2);
}

int spirk_op_cheb_first(spirk_ctx *ctx, const spirk_level *lvl, const spirk_opdesc *op, double *x1, double *x2, const double *rhs,
                        long long stride, const double *f0, const double *f1, const double *f2)
{
  return cheb_first_impl(ctx, lvl, op, x1, x2, rhs, nullptr, nullptr, stride, f0, f1, f2);
}

int spirk_op_cheb_first_diag(spirk_ctx *ctx, const spirk_level *lvl, const spirk_opdesc *op, double *x1, double *x2, const double *rhs,
                             const double *diag_mass, const double *diag_laplace, long long stride, const double *f0, const double *f1,
                             const double *f2)
{
  if (!diag_mass || !diag_laplace)
    return set_error(SPIRK_ERR_INVALID, "op_cheb_first_diag: the coefficients of the diagonal are required");
  return cheb_first_impl(ctx, lvl, op, x1, x2, rhs, diag_mass, diag_laplace, stride, f0, f1, f2);
}

int spirk_op_inverse_diagonal(spirk_ctx *ctx, const spirk_level *lvl, double *diag, double mass, double laplace)
{
  if (int e = check_level(lvl))
    return e;
  const Geo    g  = make_geo(lvl);
  const double cm = mass * std::pow(g.h, g.dim), cl = laplace * std::pow(g.h, g.dim - 2);
  SPIRK_DISPATCH_K(g.k, (k_inverse_diagonal<K><<<grid_for(ctx, g.N, 256), 256, 0, ctx->stream>>>(g, cm, cl, diag)));
  SPIRK_LAUNCH_CHECK(ctx);
  return SPIRK_OK;
}

int spirk_op_assemble_dense(spirk_ctx *ctx, const spirk_level *lvl, double mass, double laplace, double *host_matrix)
{
  if (int e = check_level(lvl))
    return e;
  const Geo g = make_geo(lvl);
  if (g.N > 4096)
    return set_error(SPIRK_ERR_INVALID, "assemble_dense: level too large");
  spirk_opdesc op;
  std::memset(&op, 0, sizeof(op));
  op.kind = SPIRK_OP_REAL, op.nb = 1, op.mass[0] = mass, op.laplace[0] = laplace;
  const size_t N = g.N;
  double      *d_e = nullptr, *d_c = nullptr;
  SPIRK_CUDA(cudaMalloc(&d_e, N * sizeof(double)));
  SPIRK_CUDA(cudaMalloc(&d_c, N * N * sizeof(double)));
  std::vector<double> cols(N * N);
  int                 st = SPIRK_OK;
  for (size_t j = 0; j < N && st == SPIRK_OK; ++j)
    {
      cudaMemsetAsync(d_e, 0, N * sizeof(double), ctx->stream);
      k_set<<<1, 1, 0, ctx->stream>>>(d_e + j, 1, 1.0);
      ctx->launches++;
      st = apply_any(ctx, g, &op, d_c + j * N, d_e, N);
    }
  if (st == SPIRK_OK)
    {
      cudaMemcpyAsync(cols.data(), d_c, N * N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
      cudaStreamSynchronize(ctx->stream);
      for (size_t i = 0; i < N; ++i)
        for (size_t j = 0; j < N; ++j)
          host_matrix[i * N + j] = cols[j * N + i];
    }
  cudaFree(d_e);
  cudaFree(d_c);
  return st;
}

// ------------------------------------------------------------------------------- transfer
// the sweeps of the owner-computes transfer (misc_kernels.cuh): extents of the arrays between the sweeps
static Sweep1D make_sweep(int d, int ex, int ey, int ez, int n_in, int ncc)
{
  Sweep1D w;
  w.d = d, w.ex = ex, w.ey = ey, w.ez = ez, w.n_in = n_in, w.ncc = ncc;
  w.N_out = (long long)ex * ey * ez;
  w.N_in  = w.N_out / (d == 0 ? ex : (d == 1 ? ey : ez)) * n_in;
  w.ec0 = 0, w.in_z0 = 0, w.out_z0 = 0;
  return w;
}

// what the two transfers share: extents, slab ranges and launch shapes
struct TransferShape
{
  Geo       g;                // fine level (maybe a z-slab)
  int       ncc, nf, ncn;     // coarse cells / fine nodes / coarse nodes per direction (whole mesh)
  int       ec0, nec;         // coarse cells of this slab in z
  int       zf_owned;         // owned fine planes
  int       zc0;              // global index of local plane 0 of the coarse vector as the caller passes it
};
static int transfer_shape(const spirk_level *lf, TransferShape &t)
{
  if (int e = check_level(lf))
    return e;
  if (lf->n_cells_1d % 2)
    return set_error(SPIRK_ERR_INVALID, "transfer: the fine level needs an even number of cells");
  t.g   = make_geo(lf);
  t.ncc = t.g.nc / 2, t.nf = t.g.n1, t.ncn = t.g.k * t.ncc + 1;
  if (t.g.col_size > 1 && (t.g.L_lo % 2 || t.g.L_hi % 2))
    return set_error(SPIRK_ERR_INVALID, "transfer: a z-slab needs an even number of cell layers");
  t.ec0 = t.g.L_lo / 2, t.nec = (t.g.L_hi - t.g.L_lo) / 2;
  t.zf_owned = (t.g.dim == 3) ? t.g.zo1 - t.g.zo0 : 1;
  t.zc0      = (t.g.col_size > 1 && !t.g.coarse_replicated) ? t.g.k * t.ec0 : 0;
  return SPIRK_OK;
}

int spirk_mg_prolongate_add(spirk_ctx *ctx, const spirk_level *lf, int nb, double *fine, long long fs, const double *coarse,
                            long long cs)
{
  TransferShape t;
  if (int e = transfer_shape(lf, t))
    return e;
  const Geo &g   = t.g;
  const int  ncc = t.ncc, nf = t.nf, ncn = t.ncn;
  if (ctx->opt_transfer_variant == 1 && g.col_size == 1)
    {
      const long long ncells = (g.dim == 3 ? (long long)ncc * ncc * ncc : (long long)ncc * ncc) * nb;
      const int       grid   = (int)std::min<long long>(ncells, (long long)ctx->n_sms * 16);
      if (g.dim == 3)
        {
          SPIRK_DISPATCH_K(g.k, (k_prolongate_add<K, 3><<<grid, 128, 0, ctx->stream>>>(g, nb, fine, fs, coarse, cs)));
        }
      else
        {
          SPIRK_DISPATCH_K(g.k, (k_prolongate_add<K, 2><<<grid, 128, 0, ctx->stream>>>(g, nb, fine, fs, coarse, cs)));
        }
      SPIRK_LAUNCH_CHECK(ctx);
      return SPIRK_OK;
    }
  // expand z, then y (one thread per coarse cell and entry of the other directions, lanes along x), then x from rows staged in
  // shared memory, adding into the fine vector.  z-slab: the coarse cells of this slab produce exactly its owned fine planes
  // (they read one coarse ghost plane above).
  constexpr int RPB = 8;
  const auto grid_1d = [&](const Sweep1D &w, int cells) {
    return dim3((unsigned)(((w.d == 1 ? w.ex * w.ez : w.ex * w.ey) + 255) / 256), (unsigned)cells, (unsigned)nb);
  };
  const dim3   blk_x(32, RPB);
  const size_t smem_x = sizeof(double) * ((size_t)RPB * ncn + (2 * g.k + 1) * (g.k + 1));
  if (g.dim == 3)
    {
      const int zf = t.zf_owned;
      Sweep1D   wz = make_sweep(2, ncn, ncn, zf, ncn, ncc), wy = make_sweep(1, ncn, nf, zf, ncn, ncc);
      wz.ec0 = t.ec0, wz.in_z0 = t.zc0, wz.out_z0 = g.zo0;
      if (int e = ensure_scratch(ctx, (size_t)nb * (wz.N_out + wy.N_out)))
        return e;
      double *t2 = ctx->d_scratch, *t1 = ctx->d_scratch + (size_t)nb * wz.N_out;
      SPIRK_DISPATCH_K(g.k, (k_prolongate_1d<K><<<grid_1d(wz, t.nec), 256, 0, ctx->stream>>>(wz, t2, wz.N_out, coarse, cs)));
      SPIRK_LAUNCH_CHECK(ctx);
      SPIRK_DISPATCH_K(g.k, (k_prolongate_1d<K><<<grid_1d(wy, ncc), 256, 0, ctx->stream>>>(wy, t1, wy.N_out, t2, wz.N_out)));
      SPIRK_LAUNCH_CHECK(ctx);
      const long long rpb = (long long)nf * zf, rows = rpb * nb;
      if (smem_x > 48 * 1024)
        SPIRK_DISPATCH_K(g.k, SPIRK_CUDA(cudaFuncSetAttribute(k_prolongate_x_add<K, RPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_x)));
      SPIRK_DISPATCH_K(g.k, (k_prolongate_x_add<K, RPB><<<grid_for(ctx, (rows + RPB - 1) / RPB, 1, 16), blk_x, smem_x, ctx->stream>>>(
                              nf, ncc, rows, rpb, fine, fs, t1, wy.N_out)));
      SPIRK_LAUNCH_CHECK(ctx);
    }
  else
    {
      const Sweep1D wy = make_sweep(1, ncn, nf, 1, ncn, ncc);
      if (int e = ensure_scratch(ctx, (size_t)nb * wy.N_out))
        return e;
      double *t1 = ctx->d_scratch;
      SPIRK_DISPATCH_K(g.k, (k_prolongate_1d<K><<<grid_1d(wy, ncc), 256, 0, ctx->stream>>>(wy, t1, wy.N_out, coarse, cs)));
      SPIRK_LAUNCH_CHECK(ctx);
      const long long rpb = nf, rows = rpb * nb;
      if (smem_x > 48 * 1024)
        SPIRK_DISPATCH_K(g.k, SPIRK_CUDA(cudaFuncSetAttribute(k_prolongate_x_add<K, RPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_x)));
      SPIRK_DISPATCH_K(g.k, (k_prolongate_x_add<K, RPB><<<grid_for(ctx, (rows + RPB - 1) / RPB, 1, 16), blk_x, smem_x, ctx->stream>>>(
                              nf, ncc, rows, rpb, fine, fs, t1, wy.N_out)));
      SPIRK_LAUNCH_CHECK(ctx);
    }
  return SPIRK_OK;
}

int spirk_mg_restrict(spirk_ctx *ctx, const spirk_level *lf, int nb, double *coarse, long long cs, const double *fine,
                      long long fs)
{
  TransferShape t;
  if (int e = transfer_shape(lf, t))
    return e;
  const Geo &g   = t.g;
  const int  ncc = t.ncc, nf = t.nf, ncn = t.ncn;
  if (ctx->opt_transfer_variant == 1 && g.col_size == 1)
    {
      // cell-based variant (atomics into a zeroed coarse vector; kept for A/B measurements)
      spirk_level lc = *lf;
      lc.n_cells_1d /= 2;
      const Geo gc = make_geo(&lc);
      for (int b = 0; b < nb; ++b)
        SPIRK_CUDA(cudaMemsetAsync(coarse + b * cs, 0, gc.N * sizeof(double), ctx->stream));
      const long long ncells = (g.dim == 3 ? (long long)ncc * ncc * ncc : (long long)ncc * ncc) * nb;
      const int       grid   = (int)std::min<long long>(ncells, (long long)ctx->n_sms * 16);
      if (g.dim == 3)
        {
          SPIRK_DISPATCH_K(g.k, (k_restrict<K, 3><<<grid, 128, 0, ctx->stream>>>(g, nb, coarse, cs, fine, fs)));
        }
      else
        {
          SPIRK_DISPATCH_K(g.k, (k_restrict<K, 2><<<grid, 128, 0, ctx->stream>>>(g, nb, coarse, cs, fine, fs)));
        }
      SPIRK_LAUNCH_CHECK(ctx);
      return SPIRK_OK;
    }
  // contract x (fine rows staged in shared memory), then y, then z: every coarse entry is written exactly once.  z-slab: x
  // and y run over the owned fine planes and the ghost planes around them (2k below, 1 above), z produces the coarse planes
  // of this slab's coarse cells.
  constexpr int RPB = 8;
  const auto grid_1d = [&](const Sweep1D &w, int cells) {
    return dim3((unsigned)(((w.d == 1 ? w.ex * w.ez : w.ex * w.ey) + 255) / 256), (unsigned)cells, (unsigned)nb);
  };
  const dim3   blk_x(32, RPB);
  const size_t smem_x = sizeof(double) * ((size_t)RPB * (nf + ncn));
  if (g.dim == 3 && g.col_size == 1 && ctx->opt_transfer_variant != 2)
    {
      // whole mesh: z first, x last.  The sweeps along y / z run at the copy bandwidth (lanes along x, no staging), the x
      // sweep at about a third of it (rows staged in shared memory): it gets the smallest array.  Measured at r = 6,
      // nb = 2: x, y, z = 2 x 106 + 36 + 19 us; z, y, x: see profiles/README.md (`transfer_variant` 2 = the former order)
      const Sweep1D wz = make_sweep(2, nf, nf, ncn, nf, ncc), wy = make_sweep(1, nf, ncn, ncn, nf, ncc);
      if (int e = ensure_scratch(ctx, (size_t)nb * (wz.N_out + wy.N_out)))
        return e;
      double         *t1 = ctx->d_scratch, *t2 = ctx->d_scratch + (size_t)nb * wz.N_out;
      const long long rpb = (long long)ncn * ncn, rows = rpb * nb;
      SPIRK_DISPATCH_K(g.k, (k_restrict_1d<K><<<grid_1d(wz, ncc), 256, 0, ctx->stream>>>(wz, t1, wz.N_out, fine, fs)));
      SPIRK_LAUNCH_CHECK(ctx);
      SPIRK_DISPATCH_K(g.k, (k_restrict_1d<K><<<grid_1d(wy, ncc), 256, 0, ctx->stream>>>(wy, t2, wy.N_out, t1, wz.N_out)));
      SPIRK_LAUNCH_CHECK(ctx);
      if (smem_x > 48 * 1024)
        SPIRK_DISPATCH_K(g.k, SPIRK_CUDA(cudaFuncSetAttribute(k_restrict_x<K, RPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_x)));
      SPIRK_DISPATCH_K(g.k, (k_restrict_x<K, RPB><<<grid_for(ctx, (rows + RPB - 1) / RPB, 1, 16), blk_x, smem_x, ctx->stream>>>(
                              nf, ncc, rows, rpb, coarse, cs, t2, wy.N_out)));
      SPIRK_LAUNCH_CHECK(ctx);
    }
  else if (g.dim == 3)
    {
      const int       zin = g.gh_lo + t.zf_owned + g.gh_hi; // fine planes in memory per block
      const Sweep1D   wx = make_sweep(0, ncn, nf, zin, nf, ncc), wy = make_sweep(1, ncn, ncn, zin, nf, ncc);
      Sweep1D         wz = make_sweep(2, ncn, ncn, ncn, zin, ncc);
      wz.ec0 = t.ec0, wz.in_z0 = g.zo0 - g.gh_lo, wz.out_z0 = t.zc0;
      if (int e = ensure_scratch(ctx, (size_t)nb * (wx.N_out + wy.N_out)))
        return e;
      double         *t1 = ctx->d_scratch, *t2 = ctx->d_scratch + (size_t)nb * wx.N_out;
      const long long rpb = (long long)nf * zin, rows = rpb * nb;
      if (smem_x > 48 * 1024) // (r >= 7: eight fine + coarse rows exceed the default dynamic shared-memory limit)
        SPIRK_DISPATCH_K(g.k, SPIRK_CUDA(cudaFuncSetAttribute(k_restrict_x<K, RPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_x)));
      SPIRK_DISPATCH_K(g.k, (k_restrict_x<K, RPB><<<grid_for(ctx, (rows + RPB - 1) / RPB, 1, 16), blk_x, smem_x, ctx->stream>>>(
                              nf, ncc, rows, rpb, t1, wx.N_out, fine - (long long)g.gh_lo * g.plane, fs)));
      SPIRK_LAUNCH_CHECK(ctx);
      SPIRK_DISPATCH_K(g.k, (k_restrict_1d<K><<<grid_1d(wy, ncc), 256, 0, ctx->stream>>>(wy, t2, wy.N_out, t1, wx.N_out)));
      SPIRK_LAUNCH_CHECK(ctx);
      SPIRK_DISPATCH_K(g.k, (k_restrict_1d<K><<<grid_1d(wz, t.nec), 256, 0, ctx->stream>>>(wz, coarse, cs, t2, wy.N_out)));
      SPIRK_LAUNCH_CHECK(ctx);
    }
  else
    {
      const Sweep1D wx = make_sweep(0, ncn, nf, 1, nf, ncc), wy = make_sweep(1, ncn, ncn, 1, nf, ncc);
      if (int e = ensure_scratch(ctx, (size_t)nb * wx.N_out))
        return e;
      double         *t1 = ctx->d_scratch;
      const long long rpb = nf, rows = rpb * nb;
      if (smem_x > 48 * 1024)
        SPIRK_DISPATCH_K(g.k, SPIRK_CUDA(cudaFuncSetAttribute(k_restrict_x<K, RPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_x)));
      SPIRK_DISPATCH_K(g.k, (k_restrict_x<K, RPB><<<grid_for(ctx, (rows + RPB - 1) / RPB, 1, 16), blk_x, smem_x, ctx->stream>>>(
                              nf, ncc, rows, rpb, t1, wx.N_out, fine, fs)));
      SPIRK_LAUNCH_CHECK(ctx);
      SPIRK_DISPATCH_K(g.k, (k_restrict_1d<K><<<grid_1d(wy, ncc), 256, 0, ctx->stream>>>(wy, coarse, cs, t1, wx.N_out)));
      SPIRK_LAUNCH_CHECK(ctx);
    }
  return SPIRK_OK;
}

int spirk_dense_matvec(spirk_ctx *ctx, int n, int nb, double *y, const double *x, long long stride, const double *matrix,
                       long long matrix_stride)
{
  if (n < 1 || n > 4096 || y == x)
    return set_error(SPIRK_ERR_INVALID, "dense_matvec: bad size or aliasing");
  k_dense_matvec<<<nb, 128, n * sizeof(double), ctx->stream>>>(n, y, x, stride, matrix, matrix_stride);
  SPIRK_LAUNCH_CHECK(ctx);
  return SPIRK_OK;
}

// ------------------------------------------------------------------------------- vectors
#define VEC_LAUNCH(kernel, n, ...)                                                       \
  do                                                                                     \
    {                                                                                    \
      if ((n) > 0)                                                                       \
        {                                                                                \
          kernel<<<grid_for(ctx, (n), 256), 256, 0, ctx->stream>>>(__VA_ARGS__);         \
          SPIRK_LAUNCH_CHECK(ctx);                                                       \
        }                                                                                \
    }                                                                                    \
  while (0)

int spirk_vec_set(spirk_ctx *ctx, double *x, long long n, double v)
{
  VEC_LAUNCH(k_set, n, x, n, v);
  return SPIRK_OK;
}
int spirk_vec_copy(spirk_ctx *ctx, double *dst, const double *src, long long n)
{
  SPIRK_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  return SPIRK_OK;
}
int spirk_vec_scale(spirk_ctx *ctx, double *x, long long n, double a)
{
  VEC_LAUNCH(k_scale, n, x, n, a);
  return SPIRK_OK;
}
int spirk_vec_axpy(spirk_ctx *ctx, double *y, double a, const double *x, long long n)
{
  VEC_LAUNCH(k_axpy, n, y, a, x, n);
  return SPIRK_OK;
}
int spirk_vec_sadd(spirk_ctx *ctx, double *y, double s, double a, const double *x, long long n)
{
  VEC_LAUNCH(k_sadd, n, y, s, a, x, n);
  return SPIRK_OK;
}
int spirk_vec_add2(spirk_ctx *ctx, double *y, double a, const double *x, double b, const double *z, long long n)
{
  VEC_LAUNCH(k_add2, n, y, a, x, b, z, n);
  return SPIRK_OK;
}
int spirk_vec_equ(spirk_ctx *ctx, double *y, double a, const double *x, long long n)
{
  VEC_LAUNCH(k_equ, n, y, a, x, n);
  return SPIRK_OK;
}
int spirk_vec_scale_pointwise(spirk_ctx *ctx, int nb, long long n, double *y, const double *d, const double *x,
                              long long stride, const double *f)
{
  if (nb < 1 || nb > SPIRK_MAX_BLOCKS)
    return set_error(SPIRK_ERR_INVALID, "scale_pointwise: nb");
  BlockFactors bf;
  for (int b = 0; b < nb; ++b)
    bf.f[b] = f[b];
  VEC_LAUNCH(k_scale_pointwise, n * nb, nb, n, y, d, x, stride, bf);
  return SPIRK_OK;
}

static inline int reduction_grid(const spirk_ctx *ctx, long long n)
{
  long long blocks = (n + RT * 4 - 1) / (RT * 4);
  return (int)std::max<long long>(1, std::min<long long>(blocks, ctx->n_sms * 8));
}

int spirk_vec_dot(spirk_ctx *ctx, const double *x, const double *y, long long n, double *host_result)
{
  const int grid = reduction_grid(ctx, n);
  k_dot<<<grid, RT, 0, ctx->stream>>>(x, y, n, ctx->d_partials);
  SPIRK_LAUNCH_CHECK(ctx);
  return finish_reduction(ctx, 1, grid, host_result);
}
int spirk_vec_add_and_dot(spirk_ctx *ctx, double *v, double a, const double *V, const double *W, long long n,
                          double *host_result)
{
  const int grid = reduction_grid(ctx, n);
  k_add_and_dot<<<grid, RT, 0, ctx->stream>>>(v, a, V, W, n, ctx->d_partials);
  SPIRK_LAUNCH_CHECK(ctx);
  return finish_reduction(ctx, 1, grid, host_result);
}
int spirk_vec_sum(spirk_ctx *ctx, const double *x, long long n, double *host_result)
{
  const int grid = reduction_grid(ctx, n);
  k_sum<<<grid, RT, 0, ctx->stream>>>(x, n, ctx->d_partials);
  SPIRK_LAUNCH_CHECK(ctx);
  return finish_reduction(ctx, 1, grid, host_result);
}

static inline dim3 reduction_grid_strided(const spirk_ctx *ctx, long long n, int nb)
{
  long long blocks = (n + RT * 4 - 1) / (RT * 4);
  return dim3((unsigned)std::max<long long>(1, std::min<long long>(blocks, std::max(1, ctx->n_partials / nb))), (unsigned)nb);
}
int spirk_vec_dot_strided(spirk_ctx *ctx, const double *x, const double *y, long long n, int nb, long long stride, double *host_result)
{
  if (nb < 1 || nb > SPIRK_MAX_BLOCKS)
    return set_error(SPIRK_ERR_INVALID, "dot_strided: nb");
  const dim3 grid = reduction_grid_strided(ctx, n, nb);
  k_dot_strided<<<grid, RT, 0, ctx->stream>>>(x, y, n, stride, ctx->d_partials);
  SPIRK_LAUNCH_CHECK(ctx);
  return finish_reduction(ctx, 1, (int)(grid.x * grid.y), host_result);
}
int spirk_vec_sum_strided(spirk_ctx *ctx, const double *x, long long n, int nb, long long stride, double *host_result)
{
  if (nb < 1 || nb > SPIRK_MAX_BLOCKS)
    return set_error(SPIRK_ERR_INVALID, "sum_strided: nb");
  const dim3 grid = reduction_grid_strided(ctx, n, nb);
  k_sum_strided<<<grid, RT, 0, ctx->stream>>>(x, n, stride, ctx->d_partials);
  SPIRK_LAUNCH_CHECK(ctx);
  return finish_reduction(ctx, 1, (int)(grid.x * grid.y), host_result);
}
int spirk_vec_add_and_dot_strided(spirk_ctx *ctx, double *v, double a, const double *V, const double *W, long long n, int nb,
                                  long long stride, double *host_result)
{
  if (nb < 1 || nb > SPIRK_MAX_BLOCKS)
    return set_error(SPIRK_ERR_INVALID, "add_and_dot_strided: nb");
  const dim3 grid = reduction_grid_strided(ctx, n, nb);
  k_add_and_dot_strided<<<grid, RT, 0, ctx->stream>>>(v, a, nullptr, V, W, n, stride, ctx->d_partials);
  SPIRK_LAUNCH_CHECK(ctx);
  return finish_reduction(ctx, 1, (int)(grid.x * grid.y), host_result);
}
int spirk_gmres_mgs_strided(spirk_ctx *ctx, double *vv, const double *const *basis, int dim, long long n, int nb, long long stride,
                            double *h, double *norm)
{
  if (dim < 1 || dim > 62 || nb < 1 || nb > SPIRK_MAX_BLOCKS)
    return set_error(SPIRK_ERR_INVALID, "gmres_mgs_strided: dim / nb");
  const dim3 grid = reduction_grid_strided(ctx, n, nb);
  const int  np   = (int)(grid.x * grid.y);
  auto       finish = [&](int slot) -> int {
    k_finish<<<1, RT, 0, ctx->stream>>>(ctx->d_partials, np, ctx->d_result + slot);
    SPIRK_LAUNCH_CHECK(ctx);
    if (ctx->reduction_comm)
      return spirk_comm_allreduce_sum(ctx, ctx->reduction_comm, ctx->d_result + slot, 1);
    return SPIRK_OK;
  };
  k_dot_strided<<<grid, RT, 0, ctx->stream>>>(vv, basis[0], n, stride, ctx->d_partials);
  SPIRK_LAUNCH_CHECK(ctx);
  if (int e = finish(0))
    return e;
  for (int i = 1; i <= dim; ++i)
    {
      k_add_and_dot_strided<<<grid, RT, 0, ctx->stream>>>(vv, 0.0, ctx->d_result + (i - 1), basis[i - 1], (i < dim) ? basis[i] : vv, n,
                                                           stride, ctx->d_partials);
      SPIRK_LAUNCH_CHECK(ctx);
      if (int e = finish(i))
        return e;
    }
  SPIRK_CUDA(cudaMemcpyAsync(ctx->h_result, ctx->d_result, (dim + 1) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  SPIRK_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < dim; ++i)
    h[i] = ctx->h_result[i];
  *norm = std::sqrt(ctx->h_result[dim]);
  return SPIRK_OK;
}

int spirk_gmres_mgs(spirk_ctx *ctx, double *vv, const double *const *basis, int dim, long long n, double *h, double *norm)
{
  if (dim < 1 || dim > 62)
    return set_error(SPIRK_ERR_INVALID, "gmres_mgs: dim");
  // h_i and the final norm stay in d_result[0 .. dim]; every update reads its coefficient from there (and, with a
  // reduction communicator attached, after the stream-ordered all-reduce of that scalar): ONE host synchronisation per sweep
  const int grid = reduction_grid(ctx, n);
  auto      finish = [&](int slot) -> int {
    k_finish<<<1, RT, 0, ctx->stream>>>(ctx->d_partials, grid, ctx->d_result + slot);
    SPIRK_LAUNCH_CHECK(ctx);
    if (ctx->reduction_comm)
      return spirk_comm_allreduce_sum(ctx, ctx->reduction_comm, ctx->d_result + slot, 1);
    return SPIRK_OK;
  };
  k_dot<<<grid, RT, 0, ctx->stream>>>(vv, basis[0], n, ctx->d_partials);
  SPIRK_LAUNCH_CHECK(ctx);
  if (int e = finish(0))
    return e;
  for (int i = 1; i <= dim; ++i)
    {
      k_sub_and_dot_dev<<<grid, RT, 0, ctx->stream>>>(vv, ctx->d_result + (i - 1), basis[i - 1], (i < dim) ? basis[i] : vv, n,
                                                      ctx->d_partials);
      SPIRK_LAUNCH_CHECK(ctx);
      if (int e = finish(i))
        return e;
    }
  SPIRK_CUDA(cudaMemcpyAsync(ctx->h_result, ctx->d_result, (dim + 1) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  SPIRK_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < dim; ++i)
    h[i] = ctx->h_result[i];
  *norm = std::sqrt(ctx->h_result[dim]);
  return SPIRK_OK;
}

int spirk_mix(spirk_ctx *ctx, int qo, int qi, double *dst, long long ds, const double *src, long long ss, long long n,
              const double *T, int add, double cutoff)
{
  if (qo < 1 || qi < 1 || qo > SPIRK_MAX_BLOCKS || qi > SPIRK_MAX_BLOCKS)
    return set_error(SPIRK_ERR_INVALID, "mix: block counts");
  if (dst == src)
    return set_error(SPIRK_ERR_INVALID, "mix: dst must not alias src");
  MixMatrix M;
  for (int i = 0; i < qo; ++i)
    for (int j = 0; j < qi; ++j)
      M.T[i * qi + j] = (std::fabs(T[i * qi + j]) > cutoff) ? T[i * qi + j] : 0.0;
  const int grid = grid_for(ctx, n, 256);
#define MIX_CASE(Q) \
  case Q: k_mix<Q><<<grid, 256, 0, ctx->stream>>>(qo, dst, ds, src, ss, n, M, add); break;
  switch (qi)
    {
      MIX_CASE(1) MIX_CASE(2) MIX_CASE(3) MIX_CASE(4) MIX_CASE(5) MIX_CASE(6) MIX_CASE(7) MIX_CASE(8)
      MIX_CASE(9) MIX_CASE(10) MIX_CASE(11) MIX_CASE(12) MIX_CASE(13) MIX_CASE(14) MIX_CASE(15) MIX_CASE(16)
    }
  SPIRK_LAUNCH_CHECK(ctx);
  return SPIRK_OK;
}

// ------------------------------------------------------------------------------- problem
static int upload_tab(spirk_ctx *ctx, const std::vector<double> &t)
{
  if (int e = ensure_tab(ctx, t.size()))
    return e;
  SPIRK_CUDA(cudaMemcpyAsync(ctx->d_tab, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  SPIRK_CUDA(cudaStreamSynchronize(ctx->stream)); // t is a stack object of the caller
  return SPIRK_OK;
}

int spirk_problem_rhs_spatial(spirk_ctx *ctx, const spirk_level *lvl, double *r)
{
  if (int e = check_level(lvl))
    return e;
  const Geo   g = make_geo(lvl);
  const Fe1D &f = g_fe[g.k];
  // separable forcing (main.cc:3523-3539): r = r1 (x) r1 [(x) r1],
  // r1_i = sum_cells sum_q phi_i(x_q) sin(2 pi x_q) w_q h
  std::vector<double> r1(g.n1, 0.0);
  for (int c = 0; c < g.nc; ++c)
    for (int q = 0; q < f.n; ++q)
      {
        const double s = std::sin(2.0 * M_PI * (c + f.xq[q]) * g.h) * f.wq[q] * g.h;
        for (int i = 0; i < f.n; ++i)
          r1[c * g.k + i] += f.B[q * f.n + i] * s;
      }
  if (int e = upload_tab(ctx, r1))
    return e;
  VEC_LAUNCH(k_outer_product, g.N, g, ctx->d_tab, 1.0, 1, r);
  return SPIRK_OK;
}

int spirk_problem_interpolate_solution(spirk_ctx *ctx, const spirk_level *lvl, double *u, double t)
{
  if (int e = check_level(lvl))
    return e;
  const Geo           g = make_geo(lvl);
  const Fe1D         &f = g_fe[g.k];
  std::vector<double> s1(g.n1);
  for (int c = 0; c < g.nc; ++c)
    for (int i = 0; i < f.n; ++i)
      s1[c * g.k + i] = std::sin(2.0 * M_PI * (c + f.nodes[i]) * g.h);
  if (int e = upload_tab(ctx, s1))
    return e;
  const double ft = (1.0 + std::sin(M_PI * t)) * std::exp(-0.5 * t);
  VEC_LAUNCH(k_outer_product, g.N, g, ctx->d_tab, ft, 0, u);
  return SPIRK_OK;
}

int spirk_problem_error_norms(spirk_ctx *ctx, const spirk_level *lvl, const double *u, double t, double *l2, double *linf)
{
  double l2sq = 0;
  if (int e = spirk_problem_error_norms_partial(ctx, lvl, u, t, &l2sq, linf))
    return e;
  *l2 = std::sqrt(l2sq);
  return SPIRK_OK;
}

int spirk_problem_error_norms_partial(spirk_ctx *ctx, const spirk_level *lvl, const double *u, double t, double *l2, double *linf)
{
  if (int e = check_level(lvl))
    return e;
  const Geo    g  = make_geo(lvl);
  const double ft = (1.0 + std::sin(M_PI * t)) * std::exp(-0.5 * t);
  SPIRK_CUDA(cudaMemsetAsync(ctx->d_result, 0, 2 * sizeof(double), ctx->stream));
  const long long ncells = (g.dim == 3) ? (long long)g.nc * g.nc * (g.L_hi - g.L_lo) : (long long)g.nc * g.nc;
  const int       grid   = (int)std::min<long long>(ncells, (long long)ctx->n_sms * 8);
  if (g.dim == 3)
    {
      SPIRK_DISPATCH_K(g.k, (k_error_norms<K, 3><<<grid, 128, 0, ctx->stream>>>(g, u, ft, ctx->d_result)));
    }
  else
    {
      SPIRK_DISPATCH_K(g.k, (k_error_norms<K, 2><<<grid, 128, 0, ctx->stream>>>(g, u, ft, ctx->d_result)));
    }
  SPIRK_LAUNCH_CHECK(ctx);
  SPIRK_CUDA(cudaMemcpyAsync(ctx->h_result, ctx->d_result, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  SPIRK_CUDA(cudaStreamSynchronize(ctx->stream));
  *l2   = ctx->h_result[0]; // the SQUARE of the L2 norm over this level's cells (a z-slab: its share)
  *linf = ctx->h_result[1];
  return SPIRK_OK;
}

int spirk_constraints_set_zero(spirk_ctx *ctx, const spirk_level *lvl, int nb, double *u, long long stride)
{
  if (int e = check_level(lvl))
    return e;
  const Geo g = make_geo(lvl);
  VEC_LAUNCH(k_set_zero_bdry, g.N * nb, g, nb, u, stride);
  return SPIRK_OK;
}

// ------------------------------------------------------------------------------- NCCL
int spirk_comm_unique_id(char *id128)
{
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId id;
  SPIRK_NCCL(nccl.GetUniqueId(&id));
  std::memcpy(id128, &id, 128);
  return SPIRK_OK;
}
int spirk_comm_create(spirk_ctx *ctx, const char *id128, int n_ranks, int rank, spirk_comm **out)
{
  SPIRK_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  std::memcpy(&id, id128, 128);
  spirk_comm *c = new spirk_comm();
  c->rank = rank, c->n_ranks = n_ranks;
  const NcclApi &api = nccl_api();
  if (!api.ok)
    {
      delete c;
      return set_error(SPIRK_ERR_COMM, api.error);
    }
  ncclResult_t r = api.CommInitRank(&c->comm, n_ranks, id, rank);
  if (r != ncclSuccess)
    {
      delete c;
      return set_error(SPIRK_ERR_COMM, std::string("ncclCommInitRank: ") + api.GetErrorString(r));
    }
  *out = c;
  return SPIRK_OK;
}
int spirk_comm_destroy(spirk_comm *c)
{
  if (c)
    {
      if (c->comm && nccl_api().ok)
        nccl_api().CommDestroy(c->comm);
      delete c;
    }
  return SPIRK_OK;
}
int spirk_comm_rank(const spirk_comm *c, int *rank, int *n)
{
  *rank = c->rank, *n = c->n_ranks;
  return SPIRK_OK;
}
int spirk_comm_allreduce_sum(spirk_ctx *ctx, spirk_comm *c, double *buf, long long n)
{
  if (c->n_ranks == 1)
    return SPIRK_OK;
  SPIRK_NCCL(nccl.AllReduce(buf, buf, n, ncclDouble, ncclSum, c->comm, ctx->stream));
  ctx->launches++;
  return SPIRK_OK;
}
int spirk_comm_split(spirk_ctx *ctx, spirk_comm *c, int color, int key, spirk_comm **out)
{
  SPIRK_CUDA(cudaSetDevice(ctx->device));
  // sizes / ranks of the new communicator: all ranks exchange (color, key)
  std::vector<int> mine = {color, key}, all((size_t)2 * c->n_ranks);
  int             *d    = nullptr;
  SPIRK_CUDA(cudaMalloc(&d, sizeof(int) * 2 * (c->n_ranks + 1)));
  SPIRK_CUDA(cudaMemcpyAsync(d, mine.data(), 2 * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  if (c->n_ranks > 1)
    SPIRK_NCCL(nccl.AllGather(d, d + 2, 2, ncclInt, c->comm, ctx->stream));
  else
    SPIRK_CUDA(cudaMemcpyAsync(d + 2, d, 2 * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
  SPIRK_CUDA(cudaMemcpyAsync(all.data(), d + 2, sizeof(int) * 2 * c->n_ranks, cudaMemcpyDeviceToHost, ctx->stream));
  SPIRK_CUDA(cudaStreamSynchronize(ctx->stream));
  SPIRK_CUDA(cudaFree(d));
  spirk_comm *s = new spirk_comm();
  for (int r = 0; r < c->n_ranks; ++r)
    if (all[2 * r] == color)
      {
        if (all[2 * r + 1] < key || (all[2 * r + 1] == key && r < c->rank))
          s->rank++;
        s->n_ranks++;
      }
  s->n_ranks -= 1; // (the struct starts at n_ranks = 1)
  if (c->n_ranks > 1)
    {
      const NcclApi &api = nccl_api();
      ncclResult_t   r   = api.CommSplit(c->comm, color, key, &s->comm, nullptr);
      if (r != ncclSuccess)
        {
          delete s;
          return set_error(SPIRK_ERR_COMM, std::string("ncclCommSplit: ") + api.GetErrorString(r));
        }
    }
  *out = s;
  return SPIRK_OK;
}
int spirk_comm_allreduce_max(spirk_ctx *ctx, spirk_comm *c, double *buf, long long n)
{
  if (c->n_ranks == 1)
    return SPIRK_OK;
  SPIRK_NCCL(nccl.AllReduce(buf, buf, n, ncclDouble, ncclMax, c->comm, ctx->stream));
  ctx->launches++;
  return SPIRK_OK;
}
int spirk_halo_exchange(spirk_ctx *ctx, spirk_comm *c, const spirk_level *lvl, int nb, double *vec, long long stride, int n_lo, int n_hi)
{
  if (int e = check_level(lvl))
    return e;
  const Geo g = make_geo(lvl);
  if (g.col_size == 1)
    return SPIRK_OK;
  if (!c || c->n_ranks != g.col_size || c->rank != g.col_rank || n_lo < 0 || n_hi < 0 || n_lo > g.gh_lo || n_hi > g.gh_hi ||
      n_lo > g.zo1 - g.zo0 || n_hi > g.zo1 - g.zo0)
    return set_error(SPIRK_ERR_INVALID, "halo_exchange: communicator / ghost depth do not match the level");
  const long long own = g.N;
  // my ghost planes below = the top n_lo owned planes of the slab below; my ghost planes above = the bottom n_hi owned planes
  // of the slab above (the top slab also owns the top plane of the domain: its top n_lo planes END at k L_hi of ... its own
  // last plane is irrelevant to a neighbour, and it has no upper neighbour)
  const bool      has_lo = g.col_rank > 0, has_hi = g.col_rank + 1 < g.col_size;
  SPIRK_NCCL(nccl.GroupStart());
  for (int b = 0; b < nb; ++b)
    {
      double *o = vec + b * stride;
      if (has_hi && n_lo > 0) // my top n_lo owned planes -> ghost planes below of the slab above
        SPIRK_NCCL(nccl.Send(o + own - (long long)n_lo * g.plane, (size_t)n_lo * g.plane, ncclDouble, c->rank + 1, c->comm, ctx->stream));
      if (has_lo && n_lo > 0)
        SPIRK_NCCL(nccl.Recv(o - (long long)n_lo * g.plane, (size_t)n_lo * g.plane, ncclDouble, c->rank - 1, c->comm, ctx->stream));
      if (has_lo && n_hi > 0) // my bottom n_hi owned planes -> ghost planes above of the slab below
        SPIRK_NCCL(nccl.Send(o, (size_t)n_hi * g.plane, ncclDouble, c->rank - 1, c->comm, ctx->stream));
      if (has_hi && n_hi > 0)
        SPIRK_NCCL(nccl.Recv(o + own, (size_t)n_hi * g.plane, ncclDouble, c->rank + 1, c->comm, ctx->stream));
    }
  SPIRK_NCCL(nccl.GroupEnd());
  ctx->launches++;
  return SPIRK_OK;
}
int spirk_comm_allgather(spirk_ctx *ctx, spirk_comm *c, double *recv, const double *send, long long n)
{
  if (c->n_ranks == 1)
    {
      if (recv != send)
        SPIRK_CUDA(cudaMemcpyAsync(recv, send, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
      return SPIRK_OK;
    }
  SPIRK_NCCL(nccl.AllGather(send, recv, n, ncclDouble, c->comm, ctx->stream));
  ctx->launches++;
  return SPIRK_OK;
}
// ---- peer-mapped exchange buffers + fused all-gather / mixing kernel
struct spirk_xbuf
{
  double               *local = nullptr;
  std::vector<double *> peer; // peer[r]: mapping of rank r's buffer (peer[rank] == local)
  long long             n = 0;
  int                   rank = 0, n_ranks = 1;
  double               *d_sync = nullptr; // 1 double for the stream-ordered rank barrier
  bool                  is_virtual = false; // member of a same-device group (spirk_xbuf_create_virtual_group)
};

int spirk_comm_xbuf_create(spirk_ctx *ctx, spirk_comm *c, long long n, spirk_xbuf **out)
{
  if (c->n_ranks > SPIRK_MAX_BLOCKS)
    return set_error(SPIRK_ERR_INVALID, "xbuf: too many ranks");
  SPIRK_CUDA(cudaSetDevice(ctx->device));
  spirk_xbuf *x = new spirk_xbuf();
  x->n = n, x->rank = c->rank, x->n_ranks = c->n_ranks;
  x->peer.assign(c->n_ranks, nullptr);
  // A failure on ONE rank must not leave the others waiting in the collective below: every rank always takes part in
  // the all-gather and ships a status byte next to its IPC handle; all ranks then agree on success or failure.
  struct Packet
  {
    cudaIpcMemHandle_t handle;
    int                ok;
  };
  Packet      mine;
  std::string why;
  std::memset(&mine, 0, sizeof(mine));
  // [0, n): this rank's published blocks; [n, 2n): result region the other ranks write into (spirk_mix_peer_a2a)
  bool ok = cudaMalloc(&x->local, (size_t)2 * n * sizeof(double)) == cudaSuccess;
  if (!ok)
    why = "cudaMalloc of the exchange buffer failed", x->local = nullptr;
  ok = ok && cudaMemsetAsync(x->local, 0, (size_t)2 * n * sizeof(double), ctx->stream) == cudaSuccess;
  ok = ok && cudaMalloc(&x->d_sync, sizeof(double)) == cudaSuccess;
  ok = ok && cudaMemsetAsync(x->d_sync, 0, sizeof(double), ctx->stream) == cudaSuccess;
  if (ok && c->n_ranks > 1 && cudaIpcGetMemHandle(&mine.handle, x->local) != cudaSuccess)
    ok = false, why = "cudaIpcGetMemHandle failed";
  cudaGetLastError();
  mine.ok = ok ? 1 : 0;
  if (ok)
    x->peer[c->rank] = x->local;
  auto fail = [&](int code, const std::string &msg) {
    for (int r = 0; r < c->n_ranks; ++r)
      if (r != c->rank && x->peer[r])
        cudaIpcCloseMemHandle(x->peer[r]);
    if (x->local)
      cudaFree(x->local);
    if (x->d_sync)
      cudaFree(x->d_sync);
    delete x;
    cudaGetLastError();
    return set_error(code, msg);
  };
  if (c->n_ranks > 1)
    {
      char *d_h = nullptr;
      if (cudaMalloc(&d_h, sizeof(Packet) * (c->n_ranks + 1)) != cudaSuccess)
        return fail(SPIRK_ERR_NOMEM, "xbuf: cudaMalloc of the handle exchange buffer failed (the other ranks may hang)");
      std::vector<Packet> all(c->n_ranks);
      const NcclApi      &api = nccl_api();
      bool                comm_ok = api.ok;
      comm_ok = comm_ok && cudaMemcpyAsync(d_h, &mine, sizeof(Packet), cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess;
      comm_ok = comm_ok && api.AllGather(d_h, d_h + sizeof(Packet), sizeof(Packet), ncclChar, c->comm, ctx->stream) == ncclSuccess;
      comm_ok = comm_ok && cudaMemcpyAsync(all.data(), d_h + sizeof(Packet), sizeof(Packet) * c->n_ranks, cudaMemcpyDeviceToHost,
                                           ctx->stream) == cudaSuccess;
      comm_ok = comm_ok && cudaStreamSynchronize(ctx->stream) == cudaSuccess;
      cudaFree(d_h);
      if (!comm_ok)
        return fail(SPIRK_ERR_COMM, "xbuf: exchange of the IPC handles failed");
      for (int r = 0; r < c->n_ranks; ++r)
        if (!all[r].ok)
          return fail(SPIRK_ERR_DEVICE, r == c->rank ? "xbuf: " + why : "xbuf: rank " + std::to_string(r) + " could not create its buffer");
      for (int r = 0; r < c->n_ranks; ++r)
        if (r != c->rank)
          {
            void *p = nullptr;
            if (cudaIpcOpenMemHandle(&p, all[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
              return fail(SPIRK_ERR_DEVICE, "xbuf: cudaIpcOpenMemHandle failed (no peer access to rank " + std::to_string(r) + ")");
            x->peer[r] = (double *)p;
          }
    }
  else if (!ok)
    return fail(SPIRK_ERR_DEVICE, "xbuf: " + why);
  *out = x;
  return SPIRK_OK;
}

int spirk_comm_xbuf_destroy(spirk_ctx *ctx, spirk_xbuf *x)
{
  if (!x)
    return SPIRK_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (!x->is_virtual)
    for (int r = 0; r < x->n_ranks; ++r)
      if (r != x->rank && x->peer[r])
        cudaIpcCloseMemHandle(x->peer[r]);
  cudaFree(x->local);
  if (x->d_sync)
    cudaFree(x->d_sync);
  delete x;
  return SPIRK_OK;
}

double *spirk_comm_xbuf_local(spirk_xbuf *x) { return x->local; }

int spirk_xbuf_create_virtual_group(spirk_ctx *ctx, int n_ranks, long long n, spirk_xbuf **out)
{
  if (n_ranks < 1 || n_ranks > SPIRK_MAX_BLOCKS || n < 1)
    return set_error(SPIRK_ERR_INVALID, "xbuf virtual group: n_ranks / n");
  SPIRK_CUDA(cudaSetDevice(ctx->device));
  std::vector<spirk_xbuf *> xs(n_ranks, nullptr);
  for (int r = 0; r < n_ranks; ++r)
    {
      spirk_xbuf *x = new spirk_xbuf();
      x->n = n, x->rank = r, x->n_ranks = n_ranks, x->is_virtual = true;
      if (cudaMalloc(&x->local, (size_t)2 * n * sizeof(double)) != cudaSuccess)
        {
          cudaGetLastError();
          delete x;
          for (spirk_xbuf *y : xs)
            if (y)
              cudaFree(y->local), delete y;
          return set_error(SPIRK_ERR_NOMEM, "xbuf virtual group: cudaMalloc failed");
        }
      cudaMemsetAsync(x->local, 0, (size_t)2 * n * sizeof(double), ctx->stream);
      xs[r] = x;
    }
  for (int r = 0; r < n_ranks; ++r)
    {
      xs[r]->peer.resize(n_ranks);
      for (int j = 0; j < n_ranks; ++j)
        xs[r]->peer[j] = xs[j]->local;
      out[r] = xs[r];
    }
  return SPIRK_OK;
}

// stream-ordered barrier over the ranks of the exchange group (a 1-double all-reduce); none for a same-device
// virtual group, whose "ranks" are ordered by the single stream they share
static int xbuf_barrier(spirk_ctx *ctx, spirk_comm *c, spirk_xbuf *x)
{
  if (c != nullptr && c->n_ranks > 1)
    SPIRK_NCCL(nccl.AllReduce(x->d_sync, x->d_sync, 1, ncclDouble, ncclSum, c->comm, ctx->stream));
  return SPIRK_OK;
}

int spirk_mix_peer(spirk_ctx *ctx, spirk_comm *c, spirk_xbuf *x, int qo, int m, double *dst, long long ds, long long n,
                   const double *T, int add, double cutoff)
{
  if (c != nullptr && (c->n_ranks != x->n_ranks || c->rank != x->rank))
    return set_error(SPIRK_ERR_INVALID, "mix_peer: communicator and exchange buffer do not match");
  const int qi = x->n_ranks * m;
  if (qo < 1 || qi > SPIRK_MAX_BLOCKS || qo > SPIRK_MAX_BLOCKS || (long long)m * n > x->n)
    return set_error(SPIRK_ERR_INVALID, "mix_peer: block counts / buffer size");
  MixMatrix M;
  for (int i = 0; i < qo; ++i)
    for (int j = 0; j < qi; ++j)
      M.T[i * qi + j] = (std::fabs(T[i * qi + j]) > cutoff) ? T[i * qi + j] : 0.0;
  PeerPtrs pp;
  for (int r = 0; r < x->n_ranks; ++r)
    pp.p[r] = x->peer[r];
  // every rank has finished writing its exchange buffer
  if (int e = xbuf_barrier(ctx, c, x))
    return e;
  const int grid = grid_for(ctx, n, 256);
#define MIXP_CASE(Q) \
  case Q: k_mix_peer<Q><<<grid, 256, 0, ctx->stream>>>(qo, m, dst, ds, pp, n, M, add); break;
  switch (qi)
    {
      MIXP_CASE(1) MIXP_CASE(2) MIXP_CASE(3) MIXP_CASE(4) MIXP_CASE(5) MIXP_CASE(6) MIXP_CASE(7) MIXP_CASE(8)
      MIXP_CASE(9) MIXP_CASE(10) MIXP_CASE(11) MIXP_CASE(12) MIXP_CASE(13) MIXP_CASE(14) MIXP_CASE(15) MIXP_CASE(16)
    }
  SPIRK_LAUNCH_CHECK(ctx);
  // ... and every rank has finished reading before a buffer may be overwritten
  return xbuf_barrier(ctx, c, x);
}

int spirk_mix_peer_a2a_contract(spirk_ctx *ctx, spirk_xbuf *x, int m, long long n, const double *T, double cutoff)
{
  const int q = x->n_ranks * m;
  if (m < 1 || q > SPIRK_MAX_BLOCKS || (long long)m * n > x->n)
    return set_error(SPIRK_ERR_INVALID, "mix_peer_a2a: block counts / buffer size");
  MixMatrix M;
  for (int i = 0; i < q; ++i)
    for (int j = 0; j < q; ++j)
      M.T[i * q + j] = (std::fabs(T[i * q + j]) > cutoff) ? T[i * q + j] : 0.0;
  PeerPtrsRW pp;
  for (int r = 0; r < x->n_ranks; ++r)
    pp.p[r] = x->peer[r];
  // this rank's chunk of every block
  const long long e0 = n * x->rank / x->n_ranks, e1 = n * (x->rank + 1) / x->n_ranks;
  const int       grid = grid_for(ctx, e1 - e0, 256);
#define MIXA_CASE(Q) \
  case Q: k_mix_a2a<Q><<<grid, 256, 0, ctx->stream>>>(m, pp, n, x->n, e0, e1, M); break;
  switch (q)
    {
      MIXA_CASE(1) MIXA_CASE(2) MIXA_CASE(3) MIXA_CASE(4) MIXA_CASE(5) MIXA_CASE(6) MIXA_CASE(7) MIXA_CASE(8)
      MIXA_CASE(9) MIXA_CASE(10) MIXA_CASE(11) MIXA_CASE(12) MIXA_CASE(13) MIXA_CASE(14) MIXA_CASE(15) MIXA_CASE(16)
    }
  SPIRK_LAUNCH_CHECK(ctx);
  return SPIRK_OK;
}

int spirk_mix_peer_a2a_finish(spirk_ctx *ctx, spirk_xbuf *x, int m, double *dst, long long ds, long long n, int add)
{
  if (m < 1 || (long long)m * n > x->n)
    return set_error(SPIRK_ERR_INVALID, "mix_peer_a2a: block counts / buffer size");
  k_mix_finish<<<grid_for(ctx, n * m, 256), 256, 0, ctx->stream>>>(m, dst, ds, x->local + x->n, n, add);
  SPIRK_LAUNCH_CHECK(ctx);
  return SPIRK_OK;
}

int spirk_mix_peer_a2a(spirk_ctx *ctx, spirk_comm *c, spirk_xbuf *x, int m, double *dst, long long ds, long long n,
                       const double *T, int add, double cutoff)
{
  if (c == nullptr || c->n_ranks != x->n_ranks || c->rank != x->rank)
    return set_error(SPIRK_ERR_INVALID, "mix_peer_a2a: communicator and exchange buffer do not match");
  // every rank has published its blocks (and is done with the previous results)
  if (int e = xbuf_barrier(ctx, c, x))
    return e;
  if (int e = spirk_mix_peer_a2a_contract(ctx, x, m, n, T, cutoff))
    return e;
  // ... and every rank's result chunks have landed here
  if (int e = xbuf_barrier(ctx, c, x))
    return e;
  return spirk_mix_peer_a2a_finish(ctx, x, m, dst, ds, n, add);
}

int spirk_ctx_set_reduction_comm(spirk_ctx *ctx, spirk_comm *comm)
{
  ctx->reduction_comm = comm;
  return SPIRK_OK;
}
}
