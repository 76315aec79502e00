/*
 * spirk_b200.h — C ABI of the B200-native device layer for the dealii-spirk hot path.
 *
 * Every entry point is `extern "C"`, takes plain pointers / sizes / PODs and returns an int
 * status (SPIRK_OK == 0); nothing throws across this boundary and no torch / C++ type appears
 * in a signature.  All `double*` vector arguments are DEVICE pointers obtained from
 * spirk_malloc unless the name says `host`.  All work is enqueued on the context's stream;
 * only functions that return a scalar to the host (dot products, norms) and spirk_ctx_sync
 * synchronise.
 *
 * The library that ships (dealii_spirk_b200/libspirk_b200.so) implements this header with
 * hand-written sm_100a CUDA kernels and fails loudly without a GPU.  A second, TEST-ONLY
 * implementation of the same header on the CPU lives in oracle/cpu_abi.cc (the "port" oracle).
 *
 * Each group cites the reference interface it replaces (paths relative to the reference root).
 *
 * Data layout: the mesh is the r-times refined unit hypercube with FE_Q(k); DoFs are numbered
 * lexicographically, index = ix + n1*(iy + n1*iz), n1 = k*2^r + 1.  A "block vector" with nb
 * blocks is nb such arrays at a common stride (>= n_dofs) from a base pointer.
 */
#ifndef SPIRK_B200_H
#define SPIRK_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPIRK_MAX_BLOCKS 16

enum
{
  SPIRK_OK              = 0,
  SPIRK_ERR_INVALID     = 1, /* bad argument */
  SPIRK_ERR_DEVICE      = 2, /* CUDA runtime error (or: no GPU) */
  SPIRK_ERR_NOMEM       = 3,
  SPIRK_ERR_UNSUPPORTED = 4, /* dim / degree / feature not compiled */
  SPIRK_ERR_COMM        = 5  /* NCCL error */
};

typedef struct spirk_ctx  spirk_ctx;  /* one per host thread / GPU: device, stream, scratch */
typedef struct spirk_comm spirk_comm; /* NCCL communicator wrapper */

/* Mesh level: replaces what MatrixFree<dim,double>::reinit(MappingQ1, dof_handler, constraints,
 * QGauss(k+1)) stores for the hypercube (operator.h:254-269, main.cc:3038-3039, 3109-3144). */
typedef struct
{
  int dim;        /* 2 or 3 */
  int degree;     /* k = 1..6 */
  int n_cells_1d; /* 2^r */
  int slab;       /* 0: the whole mesh; else SPIRK_SLAB(rank, size[, coarse_replicated]): z-slab `rank` of `size` (3-D) */
} spirk_level;
/* Spatial partition (the reference's column communicator, main.cc:3027, 3478, 3660-3698): the cell layers are split into
 * `size` equal z-slabs, n_cells_1d % size == 0.  A slab owns the node planes [k L_lo, k L_hi) (+ the top plane of the
 * domain for the last slab); its vectors are the owned planes, contiguous, and every vector pointer handed to the library
 * points at the first owned entry with SPIRK_SLAB_PAD_LO(k) ghost planes allocated below and SPIRK_SLAB_PAD_HI above
 * (a plane = (k n_cells_1d + 1)^2 doubles; blocks of a block vector each carry their own pads inside `stride`).
 * spirk_halo_exchange fills the ghost planes from the neighbouring slabs.  spirk_level_n_dofs = owned entries.
 * coarse_replicated: the next coarser level is held in full by every rank of the column (agglomerated coarse levels,
 * the analogue of create_sub_comm, preconditioner.h:287-339): the transfers address it by global plane. */
#define SPIRK_SLAB(rank, size, coarse_replicated) ((rank) | ((size) << 8) | ((coarse_replicated) ? (1 << 16) : 0))
#define SPIRK_SLAB_PAD_LO(k) (2 * (k)) /* restriction reads 2k fine planes below the owned range, the operator k */
#define SPIRK_SLAB_PAD_HI 1

/* Operator descriptor.  dst_i = laplace[i] * K src_i + M * sum_j coupling[i*nb+j] src_j on
 * unconstrained DoFs, dst_i = src_i on Dirichlet DoFs.
 *   kind REAL     : coupling ignored, diagonal mass[i]      (MassLaplaceOperatorMatrixFree::vmult,
 *                   operator.h:298-310,379-421; Batched...::vmult, operator.h:811-880)
 *   kind COUPLED  : full coupling matrix (IRK SystemMatrix main.cc:1014-1028 fused into one cell
 *                   pass; complex pair operator.h:616-665 with coupling [[lre,-lim],[lim,lre]]) */
typedef struct
{
  int    kind; /* SPIRK_OP_REAL / SPIRK_OP_COUPLED */
  int    nb;   /* number of blocks, 1..SPIRK_MAX_BLOCKS */
  double mass[SPIRK_MAX_BLOCKS];
  double laplace[SPIRK_MAX_BLOCKS];
  double coupling[SPIRK_MAX_BLOCKS * SPIRK_MAX_BLOCKS];
} spirk_opdesc;

enum
{
  SPIRK_OP_REAL    = 0,
  SPIRK_OP_COUPLED = 1
};

/* ---- library / context ------------------------------------------------------------------ */
const char *spirk_backend(void);    /* "cuda-sm_100a" for the product library */
const char *spirk_last_error(void); /* message of the last non-OK status on this thread */
int         spirk_ctx_create(spirk_ctx **ctx, int device);
int         spirk_ctx_destroy(spirk_ctx *ctx);
int         spirk_ctx_sync(spirk_ctx *ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
long long spirk_ctx_launch_count(spirk_ctx *ctx);
/* CUDA-event timing on the context's stream: begin/end return elapsed milliseconds */
int spirk_ctx_timer_begin(spirk_ctx *ctx);
int spirk_ctx_timer_end(spirk_ctx *ctx, double *ms);
/* kernel variant selection for A/B tests: 0 = default (fastest validated), see DESIGN.md */
int spirk_ctx_set_option(spirk_ctx *ctx, const char *name, int value);

/* CUDA-graph capture of a launch sequence that contains no host synchronisation (the V-cycle):
 * begin -> enqueue kernels through this API -> end returns a replayable graph.  Returns
 * SPIRK_ERR_UNSUPPORTED where capture is not available (the caller then replays eagerly). */
typedef struct spirk_graph spirk_graph;
int spirk_graph_begin(spirk_ctx *ctx);
int spirk_graph_end(spirk_ctx *ctx, spirk_graph **graph);
int spirk_graph_launch(spirk_ctx *ctx, spirk_graph *graph);
int spirk_graph_destroy(spirk_graph *graph);

/* ---- memory (replaces LinearAlgebra::distributed::Vector storage, main.cc:67-68) --------- */
int spirk_malloc(spirk_ctx *ctx, double **ptr, size_t n);
int spirk_free(spirk_ctx *ctx, double *ptr);
int spirk_copy_h2d(spirk_ctx *ctx, double *dst, const double *host_src, size_t n);
int spirk_copy_d2h(spirk_ctx *ctx, double *host_dst, const double *src, size_t n);
/* pinned host staging buffers for the end-to-end path */
int spirk_malloc_host(spirk_ctx *ctx, double **ptr, size_t n);
int spirk_free_host(spirk_ctx *ctx, double *ptr);

long long spirk_level_n_dofs(const spirk_level *lvl);

/* ---- matrix-free operator (operator.h:250-460, 529-698, 749-881) ------------------------- */
/* dst = A src */
int spirk_op_apply(spirk_ctx *ctx, const spirk_level *lvl, const spirk_opdesc *op, double *dst,
                   const double *src, long long stride);
/* dst_b = laplace[b] K v_b + mass[b] M w_b in one pass over the cells (v, w: block vectors with the same stride; Dirichlet
 * rows: dst_b = v_b).  The stage-parallel system matrix dst_s = tau K v_s + sum_j A_inv[s][j] M v_j (main.cc:1580-1592)
 * is this with w = (A_inv (x) I) v, because M commutes with the stage mixing: one cell pass instead of two. */
int spirk_op_apply_km(spirk_ctx *ctx, const spirk_level *lvl, int nb, double *dst, const double *v, const double *w,
                      long long stride, const double *laplace, const double *mass);
/* dst = rhs - A src  (Multigrid::level_v_step residual, MGSmootherPrecondition::smooth) */
int spirk_op_residual(spirk_ctx *ctx, const spirk_level *lvl, const spirk_opdesc *op, double *dst,
                      const double *rhs, const double *src, long long stride);
/* one Chebyshev iteration (deal.II PreconditionChebyshev, preconditioner.h:353-373):
 *   x_new = x + f1[b] (x - x_old) + f2[b] dinv .* (rhs - A x);  x_old == NULL means x_old = 0;
 *   dinv == NULL means "the inverse diagonal of op itself" (spirk_op_inverse_diagonal with op's coefficients; COUPLED:
 *   of block b's own term coupling[b][b] M + laplace[b] K, what ComplexMassLaplaceOperator::compute_inverse_diagonal
 *   returns, operator.h:560-575) and saves reading that vector; x_new may alias x_old, not x */
int spirk_op_cheb_step(spirk_ctx *ctx, const spirk_level *lvl, const spirk_opdesc *op,
                       double *x_new, const double *x, const double *x_old, const double *rhs,
                       const double *dinv, long long stride, const double *f1, const double *f2);
/* 1 if spirk_op_cheb_step / spirk_op_cheb_first with dinv == NULL run on a fused kernel that forms the inverse diagonal on
 * the fly for this level and operator; 0 if they would materialise it on every call (pass the stored vector then) */
int spirk_op_fuses_own_diagonal(spirk_ctx *ctx, const spirk_level *lvl, const spirk_opdesc *op);
/* The same Chebyshev iteration with dinv = the inverse diagonal of diag_mass[b] M + diag_laplace[b] K, formed on the fly:
 * the smoother of a level operator whose coefficients were changed after the multigrid set-up (the stored inverse diagonal
 * of PreconditionerGMG stems from the coefficients at reinit(), preconditioner.h:343-373; the complex batched schemes set
 * it up with the constructor defaults).  32 B per DoF where spirk_op_fuses_own_diagonal() == 1. */
int spirk_op_cheb_step_diag(spirk_ctx *ctx, const spirk_level *lvl, const spirk_opdesc *op, double *x_new, const double *x,
                            const double *x_old, const double *rhs, const double *diag_mass, const double *diag_laplace,
                            long long stride, const double *f1, const double *f2);
/* the first two Chebyshev iterates from a zero start in one pass over rhs (PreconditionChebyshev::vmult, iteration 0
 * and 1): x1 = f0[b] dinv .* rhs;  x2 = x1 + f1[b] x1 + f2[b] dinv .* (rhs - A x1), dinv = the inverse diagonal of op
 * itself (REAL operators).  24 B per DoF instead of 48 for spirk_vec_scale_pointwise + spirk_op_cheb_step. */
int spirk_op_cheb_first(spirk_ctx *ctx, const spirk_level *lvl, const spirk_opdesc *op, double *x1, double *x2,
                        const double *rhs, long long stride, const double *f0, const double *f1, const double *f2);
/* ... with dinv = the inverse diagonal of diag_mass[b] M + diag_laplace[b] K (see spirk_op_cheb_step_diag) */
int spirk_op_cheb_first_diag(spirk_ctx *ctx, const spirk_level *lvl, const spirk_opdesc *op, double *x1, double *x2,
                             const double *rhs, const double *diag_mass, const double *diag_laplace, long long stride,
                             const double *f0, const double *f1, const double *f2);
/* inverse diagonal of mass*M + laplace*K: abs(d) > 1e-10 ? 1/d : 1, Dirichlet entries 1
 * (operator.h:361-373, 560-575, 775-792) */
int spirk_op_inverse_diagonal(spirk_ctx *ctx, const spirk_level *lvl, double *diag, double mass,
                              double laplace);
/* dense matrix of mass*M + laplace*K (+ identity rows on Dirichlet DoFs) for tiny levels,
 * row-major n x n on the HOST; replaces get_system_matrix(), operator.h:331-353 */
int spirk_op_assemble_dense(spirk_ctx *ctx, const spirk_level *lvl, double mass, double laplace,
                            double *host_matrix);

/* ---- multigrid transfer (MGTwoLevelTransfer, preconditioner.h:266-282) ------------------- */
/* fine += P coarse (coarse Dirichlet DoFs read as 0); lvl_fine.n_cells_1d must be even */
int spirk_mg_prolongate_add(spirk_ctx *ctx, const spirk_level *lvl_fine, int nb, double *fine,
                            long long fine_stride, const double *coarse, long long coarse_stride);
/* coarse = P^T fine, coarse Dirichlet DoFs set to 0 */
int spirk_mg_restrict(spirk_ctx *ctx, const spirk_level *lvl_fine, int nb, double *coarse,
                      long long coarse_stride, const double *fine, long long fine_stride);
/* y_b = Minv x_b for tiny dense coarse problems (n x n row-major DEVICE matrix); replaces the
 * Trilinos-ML coarse solve on the one-cell level, preconditioner.h:375-406 */
int spirk_dense_matvec(spirk_ctx *ctx, int n, int nb, double *y, const double *x, long long stride,
                       const double *matrix, long long matrix_stride /* 0: one matrix for all blocks */);

/* ---- vector kernels (deal.II Vector ops used by SolverCG / SolverGMRES / Chebyshev and the
 *      integrators; list in SURVEY 8b "Vector API the callers use") ------------------------ */
int spirk_vec_set(spirk_ctx *ctx, double *x, long long n, double value);
int spirk_vec_copy(spirk_ctx *ctx, double *dst, const double *src, long long n);
int spirk_vec_scale(spirk_ctx *ctx, double *x, long long n, double a);                 /* x *= a */
int spirk_vec_axpy(spirk_ctx *ctx, double *y, double a, const double *x, long long n); /* y += a x */
/* y = s*y + a*x   (sadd) */
int spirk_vec_sadd(spirk_ctx *ctx, double *y, double s, double a, const double *x, long long n);
/* y += a*x + b*z  (add(a,V,b,W), main.cc:2221-2224) */
int spirk_vec_add2(spirk_ctx *ctx, double *y, double a, const double *x, double b, const double *z,
                   long long n);
/* y = a*x (equ) */
int spirk_vec_equ(spirk_ctx *ctx, double *y, double a, const double *x, long long n);
/* y_b = f[b] * d_b .* x_b  (Chebyshev iteration 0) */
int spirk_vec_scale_pointwise(spirk_ctx *ctx, int nb, long long n, double *y, const double *d,
                              const double *x, long long stride, const double *f);
/* reductions: result written to HOST (synchronises).  If a reduction communicator is attached to
 * the context (spirk_ctx_set_reduction_comm) the value is all-reduced over it first, which is
 * what ReshapedVector does over the row communicator (main.cc:237-264). */
int spirk_vec_dot(spirk_ctx *ctx, const double *x, const double *y, long long n, double *host_result);
/* v += a*V; result = v . W   (add_and_dot) */
int spirk_vec_add_and_dot(spirk_ctx *ctx, double *v, double a, const double *V, const double *W,
                          long long n, double *host_result);
int spirk_vec_sum(spirk_ctx *ctx, const double *x, long long n, double *host_result);
/* the reductions over block vectors whose nb blocks of n entries sit at `stride` > n (z-slab vectors: ghost planes between
 * the blocks are skipped) */
int spirk_vec_dot_strided(spirk_ctx *ctx, const double *x, const double *y, long long n, int nb, long long stride,
                          double *host_result);
int spirk_vec_sum_strided(spirk_ctx *ctx, const double *x, long long n, int nb, long long stride, double *host_result);
int spirk_vec_add_and_dot_strided(spirk_ctx *ctx, double *v, double a, const double *V, const double *W, long long n, int nb,
                                  long long stride, double *host_result);
int spirk_gmres_mgs_strided(spirk_ctx *ctx, double *vv, const double *const *host_basis, int dim, long long n, int nb,
                            long long stride, double *host_h, double *host_norm);
/* modified Gram-Schmidt sweep of SolverGMRES in one call:
 *   h[0] = vv.q0; h[i] = (vv -= h[i-1] q_{i-1}).q_i; norm = sqrt((vv -= h[dim-1] q_{dim-1}).vv)
 * host_basis[i] is the DEVICE pointer of basis vector i; h (dim) and norm are HOST outputs */
int spirk_gmres_mgs(spirk_ctx *ctx, double *vv, const double *const *host_basis, int dim, long long n,
                    double *host_h, double *host_norm);
/* stage mixing dst_i = [dst_i +] sum_j T[i*q_in+j] src_j, skipping |T_ij| <= cutoff
 * (main.cc:1100-1104, 1164-1168, 877-891, 1511-1529); T is a HOST row-major q_out x q_in matrix */
int spirk_mix(spirk_ctx *ctx, int q_out, int q_in, double *dst, long long dst_stride,
              const double *src, long long src_stride, long long n, const double *host_T, int add,
              double cutoff);

/* ---- problem pieces (main.cc:3213-3219, 3301-3307, 3355, 3436-3469, 3495-3602) ----------- */
/* r_i = int phi_i(x) sin(2 pi x)sin(2 pi y)[sin(2 pi z)] dx with QGauss(k+1); Dirichlet entries 0.
 * create_right_hand_side(t) = g(t) * r because the forcing is separable (main.cc:3523-3539). */
int spirk_problem_rhs_spatial(spirk_ctx *ctx, const spirk_level *lvl, double *r);
/* nodal interpolation of the analytical solution at time t (main.cc:3570-3594) */
int spirk_problem_interpolate_solution(spirk_ctx *ctx, const spirk_level *lvl, double *u, double t);
/* L2 and Linf error against the analytical solution with QGauss(k+2) (main.cc:3436-3469) */
int spirk_problem_error_norms(spirk_ctx *ctx, const spirk_level *lvl, const double *u, double t,
                              double *host_l2, double *host_linf);
/* the same over the cells of this level only (a z-slab: its cells; u needs one ghost plane above): the SQUARE of the L2
 * norm and the maximum, to be summed / maximised over the column communicator by the caller */
int spirk_problem_error_norms_partial(spirk_ctx *ctx, const spirk_level *lvl, const double *u, double t,
                                      double *host_l2_squared, double *host_linf);
/* AffineConstraints::set_zero / distribute for homogeneous Dirichlet (main.cc:3307, 3355) */
int spirk_constraints_set_zero(spirk_ctx *ctx, const spirk_level *lvl, int nb, double *u,
                               long long stride);

/* ---- communication (replaces the MPI call sites listed in SURVEY 2.3) -------------------- */
int spirk_comm_unique_id(char *id128); /* 128 bytes */
int spirk_comm_create(spirk_ctx *ctx, const char *id128, int n_ranks, int rank, spirk_comm **comm);
int spirk_comm_destroy(spirk_comm *comm);
int spirk_comm_rank(const spirk_comm *comm, int *rank, int *n_ranks);
/* in-place sum over the communicator (MPI_Allreduce, main.cc:1421-1426, 241-263) */
int spirk_comm_allreduce_sum(spirk_ctx *ctx, spirk_comm *comm, double *buf, long long n);
/* MPI_Comm_split (main.cc:311, 334, 352; row / column communicators of the rectangular process grid, 3660-3698): collective */
int spirk_comm_split(spirk_ctx *ctx, spirk_comm *comm, int color, int key, spirk_comm **out);
int spirk_comm_allreduce_max(spirk_ctx *ctx, spirk_comm *comm, double *buf, long long n);
/* Ghost planes of a z-slab vector (deal.II update_ghost_values inside cell_loop / the transfers, operator.h:301-306): fills
 * the n_lo planes below and the n_hi planes above the owned range of each of the nb blocks from the neighbouring slabs over
 * the column communicator (grouped ncclSend / ncclRecv over NVLink, stream ordered).  No-op for unpartitioned levels. */
int spirk_halo_exchange(spirk_ctx *ctx, spirk_comm *column_comm, const spirk_level *lvl, int nb, double *vec, long long stride,
                        int n_lo, int n_hi);
/* recv[r*n .. r*n+n) = send of rank r (replaces the MPI_Sendrecv_replace ring, main.cc:1465-1483) */
int spirk_comm_allgather(spirk_ctx *ctx, spirk_comm *comm, double *recv, const double *send,
                         long long n);
/* Peer-mapped exchange buffers over NVLink / NVSwitch (collective): every rank allocates n doubles and
 * maps the buffers of all other ranks (CUDA IPC), the analogue of the reference's MPI-3 shared-memory
 * window (main.cc:1327-1331, 229-235, `UseSharedMemory`). */
typedef struct spirk_xbuf spirk_xbuf;
int     spirk_comm_xbuf_create(spirk_ctx *ctx, spirk_comm *comm, long long n, spirk_xbuf **xbuf);
int     spirk_comm_xbuf_destroy(spirk_ctx *ctx, spirk_xbuf *xbuf);
double *spirk_comm_xbuf_local(spirk_xbuf *xbuf); /* this rank's buffer (device pointer) */
/* Fused all-gather + stage mixing in ONE kernel over peer memory (collective):
 *   dst_i = [dst_i +] sum_j T[i*q+j] * X_j,  i < q_out,  q = n_ranks * m_per_rank,
 * where X_j is block (j % m_per_rank) (stride n) of rank (j / m_per_rank)'s exchange buffer, read
 * directly over NVLink (peer loads) — the `use_sm` variant of perform_basis_change,
 * main.cc:1506-1533.  Ranks are synchronised on the stream before and after the peer reads, so the
 * exchange buffer may be overwritten as soon as the call returns (in stream order). */
int spirk_mix_peer(spirk_ctx *ctx, spirk_comm *comm, spirk_xbuf *xbuf, int q_out, int m_per_rank, double *dst,
                   long long dst_stride, long long n, const double *host_T, int add, double cutoff);
/* The same mixing as an all-to-all (collective): host_T is the FULL q x q matrix (row-major); this rank contracts its
 * 1/n_ranks chunk of every block for all q outputs and writes each output chunk straight into the owner's result
 * region of the exchange buffer, then dst_i = [dst_i +] result_i for its own m_per_rank outputs.  2 (R-1)/R n
 * doubles cross NVLink per rank instead of (R-1) m n. */
int spirk_mix_peer_a2a(spirk_ctx *ctx, spirk_comm *comm, spirk_xbuf *xbuf, int m_per_rank, double *dst, long long dst_stride,
                       long long n, const double *host_T, int add, double cutoff);
/* The two halves of spirk_mix_peer_a2a without the rank barriers (the caller orders them): contract = this rank's chunk of
 * every block for all q outputs, written into the owners' result regions; finish = dst_i = [dst_i +] result_i. */
int spirk_mix_peer_a2a_contract(spirk_ctx *ctx, spirk_xbuf *xbuf, int m_per_rank, long long n, const double *host_T, double cutoff);
int spirk_mix_peer_a2a_finish(spirk_ctx *ctx, spirk_xbuf *xbuf, int m_per_rank, double *dst, long long dst_stride, long long n,
                              int add);
/* R exchange buffers on ONE device that see each other as peers: the ranks of a stage group emulated on a single GPU
 * (B200_PROFILING.md: with fewer GPUs than ranks, run all ranks' data through the same kernels on one device).  Used by
 * the single-GPU tests of the peer mixing kernels and by `spirk` runs with more stage ranks than devices; pass
 * comm == NULL to spirk_mix_peer (rank and group size are the buffer's; stream order replaces the rank barrier). */
int spirk_xbuf_create_virtual_group(spirk_ctx *ctx, int n_ranks, long long n, spirk_xbuf **xbufs /* n_ranks entries */);
/* attach / detach (NULL) the communicator over which dot products are summed */
int spirk_ctx_set_reduction_comm(spirk_ctx *ctx, spirk_comm *comm);

#ifdef __cplusplus
}
#endif
#endif /* SPIRK_B200_H */
