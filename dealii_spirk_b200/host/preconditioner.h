// Host mirror of the reference's include/preconditioner.h: PreconditionerBase and the geometric
// multigrid PreconditionerGMG (ref preconditioner.h:159-174, 219-501).  deal.II's Multigrid,
// PreconditionChebyshev, MGSmootherPrecondition and MGTransferGlobalCoarsening are restated per
// SURVEY Appendix A7-A10; all level work (fused Chebyshev steps, fused residuals, transfers,
// coarse solve) runs in CUDA kernels behind the C ABI, and one V-cycle is replayed as a CUDA graph.
// PreconditionerAMG (Trilinos ML on an assembled matrix, ref preconditioner.h:176-215) is out of
// scope.
#pragma once
#include <cstdlib>
#include <functional>
#include <memory>

#include "operator.h"
#include "solvers.h"

namespace spirk_host
{
  template <typename VectorType>
  class PreconditionerBase
  {
  public:
    virtual ~PreconditionerBase()                                        = default;
    virtual void reinit() const                                          = 0;
    virtual void vmult(VectorType &dst, const VectorType &src) const     = 0;
    virtual std::unique_ptr<const PreconditionerBase<VectorType>> clone() const = 0;
  };

  struct PreconditionerGMGAdditionalData
  {
    double       smoothing_range               = 20;
    unsigned int smoothing_degree              = 5;
    unsigned int smoothing_eig_cg_n_iterations = 20;

    unsigned int coarse_grid_smoother_sweeps = 1;
    unsigned int coarse_grid_n_cycles        = 1;
    std::string  coarse_grid_smoother_type   = "ILU";

    unsigned int coarse_grid_maxiter = 1000;
    double       coarse_grid_abstol  = 1e-20;
    double       coarse_grid_reltol  = 1e-4;
  };

  namespace internal
  {
    // largest / smallest eigenvalue of a symmetric tridiagonal matrix by Sturm bisection
    inline void tridiagonal_extreme_eigenvalues(const std::vector<double> &d, const std::vector<double> &e, double &lmin, double &lmax)
    {
      const int n = d.size();
      for (int i = 0; i < n; ++i)
        if (std::isnan(d[i]) || (i + 1 < n && std::isnan(e[i])))
          throw Error("Chebyshev eigenvalue estimate: Lanczos coefficients are NaN (operator not positive definite)");
      double lo = d[0], hi = d[0];
      for (int i = 0; i < n; ++i)
        {
          const double r = (i > 0 ? std::fabs(e[i - 1]) : 0.0) + (i + 1 < n ? std::fabs(e[i]) : 0.0);
          lo = std::min(lo, d[i] - r), hi = std::max(hi, d[i] + r);
        }
      auto count_below = [&](double x) { // number of eigenvalues < x
        int    c = 0;
        double q = d[0] - x;
        if (q < 0)
          ++c;
        for (int i = 1; i < n; ++i)
          {
            if (q == 0)
              q = 1e-300;
            q = d[i] - x - e[i - 1] * e[i - 1] / q;
            if (q < 0)
              ++c;
          }
        return c;
      };
      auto kth = [&](int k) { // k-th smallest, 0-based
        double a = lo, b = hi;
        for (int it = 0; it < 200 && b - a > 4e-16 * std::max(std::fabs(a), std::fabs(b)); ++it)
          {
            const double m = 0.5 * (a + b);
            if (count_below(m) > k)
              b = m;
            else
              a = m;
          }
        return 0.5 * (a + b);
      };
      lmin = kth(0), lmax = kth(n - 1);
    }

    // in-place inversion of a small dense matrix (Gauss-Jordan with partial pivoting)
    inline void invert_dense(std::vector<double> &A, int n)
    {
      std::vector<double> I((size_t)n * n, 0.0);
      for (int i = 0; i < n; ++i)
        I[(size_t)i * n + i] = 1.0;
      for (int c = 0; c < n; ++c)
        {
          int p = c;
          for (int r = c + 1; r < n; ++r)
            if (std::fabs(A[(size_t)r * n + c]) > std::fabs(A[(size_t)p * n + c]))
              p = r;
          if (A[(size_t)p * n + c] == 0.0)
            throw Error("coarse matrix is singular");
          if (p != c)
            for (int j = 0; j < n; ++j)
              std::swap(A[(size_t)p * n + j], A[(size_t)c * n + j]), std::swap(I[(size_t)p * n + j], I[(size_t)c * n + j]);
          const double inv = 1.0 / A[(size_t)c * n + c];
          for (int j = 0; j < n; ++j)
            A[(size_t)c * n + j] *= inv, I[(size_t)c * n + j] *= inv;
          for (int r = 0; r < n; ++r)
            if (r != c)
              {
                const double f = A[(size_t)r * n + c];
                if (f != 0.0)
                  for (int j = 0; j < n; ++j)
                    A[(size_t)r * n + j] -= f * A[(size_t)c * n + j], I[(size_t)r * n + j] -= f * I[(size_t)c * n + j];
              }
        }
      A.swap(I);
    }

    // DiagonalMatrix<VectorType>::vmult as the CG preconditioner of the eigenvalue estimate
    struct DiagonalPreconditioner
    {
      const Vector &dinv;
      void          vmult(Vector &dst, const Vector &src) const
      {
        std::vector<double> one(dinv.n_blocks(), 1.0);
        SPIRK_CHECK(spirk_vec_scale_pointwise(dst.ctx(), dinv.n_blocks(), dinv.block_size(), dst.data(), dinv.data(), src.data(),
                                              dinv.stride(), one.data()));
      }
    };

    struct LevelOperatorRef
    {
      const LevelOperatorBase *op;
      void                     vmult(Vector &dst, const Vector &src) const { op->vmult(dst, src); }
    };

    // Smoother + work vectors of one multigrid level for nb blocks
    struct MGLevel
    {
      spirk_level         level{};
      MatrixFree          mf;          // the level's mesh (a z-slab of it on partitioned levels) and its column communicator
      long long           n = 0;      // locally owned entries per block
      long long           stride = 0; // distance between the blocks of the work vectors (ghost planes on z-slab levels)
      Vector              dinv;
      std::vector<double> theta, delta; // per block
      std::vector<double> max_eigenvalue, min_eigenvalue;
      // coefficients (mass, laplace) the inverse diagonal of block b was computed with; when they
      // equal the live operator's, the smoother lets the kernel form D^-1 on the fly (dinv == NULL)
      std::vector<double> dinv_mass, dinv_laplace;
      unsigned int        cg_iterations = 0;
      Vector              defect, solution, t, d, tmp;
    };

    // The multigrid algorithm for nb blocks; operators enter only through C-ABI descriptors.
    class MultigridCore
    {
    public:
      Device                                    *dev = nullptr;
      int                                        nb  = 1;
      std::vector<MGLevel>                       levels;
      std::function<spirk_opdesc(unsigned int)>  opdesc; // live descriptor of level l
      unsigned int                               degree        = 5;
      bool                                       coarse_exact  = true;
      Vector                                     coarse_inverse; // nb dense n0 x n0 matrices
      bool                                       use_graph     = true;

      ~MultigridCore()
      {
        if (graph)
          spirk_graph_destroy(graph);
      }

      void allocate_work_vectors()
      {
        for (auto &L : levels)
          {
            for (Vector *v : {&L.defect, &L.solution, &L.t, &L.d, &L.tmp})
              v->reinit(*dev, L.n, nb);
            L.stride = L.defect.stride();
            if (L.mf.partitioned())
              use_graph = false; // (the halo exchanges are NCCL calls; the V-cycle of a partitioned hierarchy runs eagerly)
          }
      }

      // PreconditionChebyshev::vmult on level l (zero initial guess), result in x (A7)
      void smooth_zero_start(unsigned int l, Vector &x, const Vector &b)
      {
        MGLevel            &L = levels[l];
        std::vector<double> f(nb), f1(nb), f2(nb);
        for (int i = 0; i < nb; ++i)
          f[i] = 1.0 / L.theta[i];
        const spirk_opdesc op = opdesc(l);
        // the kernel forms D^-1 on the fly only where that is fused into the cell pass (elsewhere it would be recomputed
        // into scratch memory on every call: pass the stored diagonal)
        // (the stored inverse diagonal is the diagonal of dinv_mass[i] M + dinv_laplace[i] K, the operator's coefficients at
        // reinit(); the kernel forms exactly that per node class, whatever the live coefficients are - SURVEY 2.4(9))
        const bool own_dinv = (int)L.dinv_mass.size() == nb && spirk_op_fuses_own_diagonal(dev->ctx(), &L.level, &op);
        std::vector<double> rhok(nb), sigma(nb);
        bool                any = false;
        for (int i = 0; i < nb; ++i)
          {
            rhok[i] = L.delta[i] / L.theta[i], sigma[i] = L.theta[i] / L.delta[i];
            any = any || std::fabs(L.delta[i]) >= 1e-40;
          }
        // iterations 0 and 1 in one pass over b when the smoother's diagonal is the level operator's own
        const bool fuse_first = own_dinv && degree >= 2 && any && fuse_first_iterations();
        if (!fuse_first)
          SPIRK_CHECK(spirk_vec_scale_pointwise(dev->ctx(), nb, L.n, x.data(), L.dinv.data(), b.data(), L.stride, f.data()));
        if (degree < 2 || !any)
          return;
        double *cur = x.data(), *old = L.tmp.data();
        for (unsigned int k = 0; k < degree - 1; ++k)
          {
            for (int i = 0; i < nb; ++i)
              {
                const double rhokp = 1. / (2. * sigma[i] - rhok[i]);
                f1[i] = rhokp * rhok[i], f2[i] = 2. * rhokp / L.delta[i];
                rhok[i] = rhokp;
              }
            // x_new overwrites the x_old buffer (deal.II swaps solution / solution_old)
            // the vector the cell operator is applied to needs its ghost planes (z-slab levels)
            const int kd = L.level.degree;
            if (k == 0 && fuse_first)
              {
                L.mf.exchange_ghosts(const_cast<double *>(b.data()), nb, L.stride, kd, 1);
                SPIRK_CHECK(spirk_op_cheb_first_diag(dev->ctx(), &L.level, &op, cur, old, b.data(), L.dinv_mass.data(),
                                                     L.dinv_laplace.data(), L.stride, f.data(), f1.data(), f2.data()));
              }
            else
              {
                L.mf.exchange_ghosts(cur, nb, L.stride, kd, 1);
                if (own_dinv)
                  SPIRK_CHECK(spirk_op_cheb_step_diag(dev->ctx(), &L.level, &op, old, cur, k == 0 ? nullptr : old, b.data(),
                                                      L.dinv_mass.data(), L.dinv_laplace.data(), L.stride, f1.data(), f2.data()));
                else
                  SPIRK_CHECK(spirk_op_cheb_step(dev->ctx(), &L.level, &op, old, cur, k == 0 ? nullptr : old, b.data(),
                                                 L.dinv.data(), L.stride, f1.data(), f2.data()));
              }
            std::swap(cur, old);
          }
        if (cur != x.data()) // odd number of steps: result sits in the tmp buffer
          SPIRK_CHECK(spirk_vec_copy(dev->ctx(), x.data(), cur, (long long)(nb - 1) * L.stride + L.n));
      }

      static bool fuse_first_iterations()
      {
        static int v = -1;
        if (v < 0)
          {
            const char *e = std::getenv("SPIRK_FUSE_CHEB_FIRST"); // 0: scale + Chebyshev step as two kernels
            v             = (e && std::atoi(e) == 0) ? 0 : 1;
          }
        return v == 1;
      }

      void coarse_solve()
      {
        MGLevel &L = levels[0];
        if (coarse_exact)
          SPIRK_CHECK(spirk_dense_matvec(dev->ctx(), (int)L.n, nb, L.solution.data(), L.defect.data(), L.stride, coarse_inverse.data(),
                                         L.n * L.n));
        else
          smooth_zero_start(0, L.solution, L.defect); // MGCoarseGridApplyPreconditioner on the smoother
      }

      // Multigrid::level_v_step (A8)
      void level_v_step(unsigned int l)
      {
        if (l == 0)
          {
            coarse_solve();
            return;
          }
        MGLevel           &L  = levels[l];
        MGLevel           &Lc = levels[l - 1];
        const spirk_opdesc op = opdesc(l);
        const int kd = L.level.degree;
        smooth_zero_start(l, L.solution, L.defect);
        L.mf.exchange_ghosts_for_operator(L.solution);
        SPIRK_CHECK(spirk_op_residual(dev->ctx(), &L.level, &op, L.t.data(), L.defect.data(), L.solution.data(), L.stride));
        // restriction reads 2k fine planes below and one above the owned range; a replicated coarse level receives the planes
        // of every slab (all-gather over the column communicator; its top plane is Dirichlet)
        L.mf.exchange_ghosts(L.t, SPIRK_SLAB_PAD_LO(kd), SPIRK_SLAB_PAD_HI);
        SPIRK_CHECK(spirk_mg_restrict(dev->ctx(), &L.level, nb, Lc.defect.data(), Lc.stride, L.t.data(), L.stride));
        if (L.mf.partitioned() && !Lc.mf.partitioned())
          gather_replicated(L, Lc, Lc.defect);
        level_v_step(l - 1);
        if (Lc.mf.partitioned())
          Lc.mf.exchange_ghosts(Lc.solution, 0, 1); // prolongation reads the coarse plane on top of the slab
        SPIRK_CHECK(spirk_mg_prolongate_add(dev->ctx(), &L.level, nb, L.solution.data(), L.stride, Lc.solution.data(), Lc.stride));
        // post-smoothing: x += S (b - A x)
        L.mf.exchange_ghosts_for_operator(L.solution);
        SPIRK_CHECK(spirk_op_residual(dev->ctx(), &L.level, &op, L.t.data(), L.defect.data(), L.solution.data(), L.stride));
        smooth_zero_start(l, L.d, L.t);
        L.solution.add(1.0, L.d);
      }

      // The coarse level below a z-slab level is held in full by every rank of the column (agglomerated coarse levels:
      // every rank repeats the small coarse-level work instead of idling, the analogue of create_sub_comm,
      // preconditioner.h:287-339).  After the restriction every rank holds the coarse planes of its own slab: all-gather.
      void gather_replicated(const MGLevel &L, const MGLevel &Lc, Vector &v)
      {
        const int       col_size = (L.level.slab >> 8) & 0xff, col_rank = L.level.slab & 0xff;
        const int       n1c = Lc.mf.n1(), kd = Lc.level.degree;
        const long long plane = (long long)n1c * n1c, count = (long long)kd * (Lc.level.n_cells_1d / col_size) * plane;
        for (int b = 0; b < nb; ++b)
          {
            double *base = v.data() + b * v.stride();
            SPIRK_CHECK(spirk_comm_allgather(dev->ctx(), L.mf.column_comm, base, base + col_rank * count, count));
          }
        SPIRK_CHECK(spirk_constraints_set_zero(dev->ctx(), &Lc.level, nb, v.data(), v.stride()));
      }

      bool same_descriptors(const std::vector<spirk_opdesc> &a) const
      {
        if (a.size() != graph_descs.size())
          return false;
        for (size_t i = 0; i < a.size(); ++i)
          if (std::memcmp(&a[i], &graph_descs[i], sizeof(spirk_opdesc)) != 0)
            return false;
        return true;
      }

      // PreconditionMG::vmult: copy_to_mg, one V-cycle, copy_from_mg
      void vmult(Vector &dst, const Vector &src)
      {
        const unsigned int top = levels.size() - 1;
        levels[top].defect     = src;
        if (use_graph && !graph_unsupported)
          {
            std::vector<spirk_opdesc> now(levels.size());
            for (unsigned int l = 0; l < levels.size(); ++l)
              now[l] = opdesc(l);
            if (graph && !same_descriptors(now))
              {
                spirk_graph_destroy(graph);
                graph = nullptr, warm = false;
              }
            if (!graph)
              {
                if (!warm)
                  { // first call eagerly: lets the library size its scratch buffers outside capture
                    level_v_step(top);
                    warm = true;
                    dst  = levels[top].solution;
                    return;
                  }
                const int st = spirk_graph_begin(dev->ctx());
                if (st == SPIRK_ERR_UNSUPPORTED)
                  graph_unsupported = true;
                else
                  {
                    check(st, "spirk_graph_begin");
                    // whatever goes wrong inside the capture (a launch that cannot be captured, an allocation): end the
                    // capture so that the stream is usable again, and run this and all later V-cycles eagerly
                    bool captured = false;
                    try
                      {
                        level_v_step(top);
                        captured = true;
                      }
                    catch (const Error &)
                      {}
                    const int st_end = spirk_graph_end(dev->ctx(), &graph);
                    if (captured && st_end == SPIRK_OK)
                      graph_descs = now;
                    else
                      {
                        if (st_end == SPIRK_OK && graph)
                          spirk_graph_destroy(graph);
                        graph = nullptr, graph_unsupported = true;
                      }
                  }
              }
            if (graph)
              {
                SPIRK_CHECK(spirk_graph_launch(dev->ctx(), graph));
                dst = levels[top].solution;
                return;
              }
          }
        level_v_step(top);
        dst = levels[top].solution;
      }

    private:
      spirk_graph              *graph = nullptr;
      std::vector<spirk_opdesc> graph_descs;
      bool                      warm = false, graph_unsupported = false;
    };
  } // namespace internal

  // PreconditionerGMG<dim, LevelMatrixType, VectorType>.  LevelMatrixType == MassLaplaceOperator gives
  // the scalar V-cycle with an exact level-0 solve; block level operators (Batched / Complex) give the
  // block V-cycle whose Chebyshev eigenvalue range is estimated once for all blocks and whose coarse
  // solve is one sweep of the level-0 smoother (ref preconditioner.h:375-413).
  template <int dim, typename LevelMatrixType, typename VectorType = Vector, typename VectorTypeScalar = Vector>
  class PreconditionerGMG : public PreconditionerBase<VectorType>
  {
    static const bool working_on_block_vector = !std::is_same<LevelMatrixType, MassLaplaceOperator>::value;

  public:
    using LevelOperators = std::vector<std::shared_ptr<const LevelMatrixType>>;

    // (dof_handler, mg_dof_handlers, mg_constraints) of the reference are implied by the level operators
    PreconditionerGMG(const LevelOperators &mg_operators)
      : mg_operators(mg_operators)
      , min_level(0)
      , max_level(mg_operators.size() - 1)
    {}

    void reinit() const override
    {
      PreconditionerGMGAdditionalData additional_data;
      core.reset(new internal::MultigridCore());
      auto &c   = *core;
      c.dev     = mg_operators[0]->get_matrix_free().device;
      c.nb      = mg_operators[0]->n_blocks();
      c.degree  = additional_data.smoothing_degree;
      c.levels.resize(max_level + 1);
      const LevelOperators ops = mg_operators;
      c.opdesc = [ops](unsigned int l) { return ops[l]->descriptor(); };
      for (unsigned int level = min_level; level <= max_level; ++level)
        {
          auto &L = c.levels[level];
          L.mf    = mg_operators[level]->get_matrix_free();
          L.level = L.mf.level;
          L.n     = L.mf.n_dofs();
          mg_operators[level]->compute_inverse_diagonal(L.dinv);
          const spirk_opdesc d = mg_operators[level]->descriptor();
          if (d.kind == SPIRK_OP_REAL)
            {
              L.dinv_mass.assign(d.mass, d.mass + d.nb);
              L.dinv_laplace.assign(d.laplace, d.laplace + d.nb);
            }
          else if (d.nb == 2 && d.coupling[0] == d.coupling[3] && d.laplace[0] == d.laplace[1])
            {
              // complex pair: compute_inverse_diagonal is the diagonal of lambda_re M + tau K in both blocks
              // (operator.h:560-575) = the diagonal of each block's own term; remembered so that the smoother can let the
              // kernel form it on the fly while the live coefficients still equal the ones of this set-up
              L.dinv_mass.assign({d.coupling[0], d.coupling[3]});
              L.dinv_laplace.assign(d.laplace, d.laplace + d.nb);
            }
        }
      for (unsigned int level = min_level; level <= max_level; ++level)
        estimate_eigenvalues(c, level, additional_data);
      c.coarse_exact = !working_on_block_vector;
      if (c.coarse_exact)
        setup_coarse_inverse(c);
      c.allocate_work_vectors();
    }

    void vmult(VectorType &dst, const VectorType &src) const override
    {
      if (!core)
        throw Error("PreconditionerGMG::vmult called before reinit() (ExcInternalError)");
      core->vmult(dst, src);
    }

    std::unique_ptr<const PreconditionerBase<VectorType>> clone() const override
    {
      return std::make_unique<PreconditionerGMG<dim, LevelMatrixType, VectorType, VectorTypeScalar>>(mg_operators);
    }

    // read access for statistics / tests and for stage batching (merge of several scalar clones)
    const internal::MultigridCore &get_core() const { return *core; }
    internal::MultigridCore       &get_core() { return *core; }
    const LevelOperators          &get_level_operators() const { return mg_operators; }

  private:
    // deal.II PreconditionChebyshev::estimate_eigenvalues (A7)
    void estimate_eigenvalues(internal::MultigridCore &c, unsigned int level, const PreconditionerGMGAdditionalData &data) const
    {
      auto        &L  = c.levels[level];
      const auto  *op = mg_operators[level].get();
      Vector       rhs, sol;
      op->initialize_block_vector(rhs);
      op->initialize_block_vector(sol);
      // set_initial_guess: (global index % 11) minus its mean, block by block
      {
        // over the WHOLE level; a z-slab holds the entries [first, first + n) of it
        const long long n1 = L.mf.n1(), plane = (L.level.dim == 3) ? n1 * n1 : n1;
        const long long n_global = plane * n1;
        const int       col_size = (L.level.slab >> 8) & 0xff, col_rank = L.level.slab & 0xff;
        const long long first    = (col_size > 1) ? (long long)L.level.degree * (L.level.n_cells_1d * col_rank / col_size) * plane : 0;
        const long long rem      = n_global % 11;
        const double    sum      = 55.0 * (double)(n_global / 11) + 0.5 * (double)(rem * (rem - 1));
        const double    mean     = sum / (double)n_global;
        std::vector<double> g((size_t)L.n);
        for (long long i = 0; i < L.n; ++i)
          g[i] = (double)((first + i) % 11) - mean;
        for (int b = 0; b < c.nb; ++b)
          rhs.block(b).copy_from_host(g.data());
      }
      ReductionControl control(data.smoothing_eig_cg_n_iterations, std::sqrt(std::numeric_limits<double>::epsilon()), 1e-10);
      SolverCG         solver(control);
      solver.collect_lanczos = true;
      try
        {
          solver.solve(internal::LevelOperatorRef{op}, sol, rhs, internal::DiagonalPreconditioner{L.dinv});
        }
      catch (const SolverControl::NoConvergence &)
        {}
      double min_ev = 1, max_ev = 1;
      if (!solver.lanczos_diagonal.empty())
        {
          const size_t        n = solver.lanczos_diagonal.size();
          std::vector<double> e(solver.lanczos_offdiagonal.begin(), solver.lanczos_offdiagonal.begin() + (n - 1));
          double              lmin, lmax;
          internal::tridiagonal_extreme_eigenvalues(solver.lanczos_diagonal, e, lmin, lmax);
          min_ev = lmin;
          max_ev = 1.2 * lmax;
        }
      L.cg_iterations = control.last_step();
      const double alpha = (data.smoothing_range > 1. ? max_ev / data.smoothing_range : std::min(0.9 * max_ev, min_ev));
      L.theta.assign(c.nb, (max_ev + alpha) * 0.5);
      L.delta.assign(c.nb, (max_ev - alpha) * 0.5);
      L.max_eigenvalue.assign(c.nb, max_ev);
      L.min_eigenvalue.assign(c.nb, min_ev);
    }

    // exact level-0 solve: dense inverse of the interior block, identity on Dirichlet rows (A10)
    void setup_coarse_inverse(internal::MultigridCore &c) const
    {
      auto               &L0 = c.levels[0];
      const long long     n  = L0.n;
      const auto          A  = get_scalar_system_matrix();
      const int           n1 = L0.level.degree * L0.level.n_cells_1d + 1;
      std::vector<int>    interior;
      for (long long i = 0; i < n; ++i)
        {
          const int  ix = i % n1, iy = (i / n1) % n1, iz = (dim == 3) ? i / ((long long)n1 * n1) : 1;
          const bool bd = ix == 0 || ix == n1 - 1 || iy == 0 || iy == n1 - 1 || (dim == 3 && (iz == 0 || iz == n1 - 1));
          if (!bd)
            interior.push_back((int)i);
        }
      const int           ni = interior.size();
      std::vector<double> Ai((size_t)ni * ni);
      for (int r = 0; r < ni; ++r)
        for (int s = 0; s < ni; ++s)
          Ai[(size_t)r * ni + s] = A[(size_t)interior[r] * n + interior[s]];
      if (ni > 0)
        internal::invert_dense(Ai, ni);
      std::vector<double> full((size_t)n * n, 0.0);
      for (long long i = 0; i < n; ++i)
        full[(size_t)i * n + i] = 1.0;
      for (int r = 0; r < ni; ++r)
        for (int s = 0; s < ni; ++s)
          full[(size_t)interior[r] * n + interior[s]] = Ai[(size_t)r * ni + s];
      c.coarse_inverse.reinit(*c.dev, n * n, 1, true);
      c.coarse_inverse.copy_from_host(full.data());
    }

    std::vector<double> get_scalar_system_matrix() const
    {
      if constexpr (!working_on_block_vector)
        return mg_operators[0]->get_system_matrix();
      else
        return {};
    }

    const LevelOperators mg_operators;
    const unsigned int   min_level;
    const unsigned int   max_level;

    mutable std::unique_ptr<internal::MultigridCore> core;
  };

  // Stage batching: q scalar PreconditionerGMG clones (one per RK stage, each set up with its own
  // (d_i, tau), ref main.cc:1083-1089) applied in lock-step as ONE nb = q block V-cycle with
  // per-block coefficients, Chebyshev parameters and coarse inverses.  Numerically this is the q
  // independent V-cycles of the `irk` / `spirk` paths; on the GPU it is one set of batched kernels.
  template <int dim>
  class StageBatchedGMG
  {
  public:
    using ScalarGMG = PreconditionerGMG<dim, MassLaplaceOperator, Vector>;

    StageBatchedGMG(const std::vector<const ScalarGMG *> &clones, const std::vector<double> &mass, const std::vector<double> &laplace)
    {
      const int   nb = clones.size();
      const auto &c0 = clones[0]->get_core();
      core.dev = c0.dev, core.nb = nb, core.degree = c0.degree, core.coarse_exact = true;
      core.levels.resize(c0.levels.size());
      for (unsigned int l = 0; l < c0.levels.size(); ++l)
        {
          auto &L = core.levels[l];
          L.level = c0.levels[l].level, L.n = c0.levels[l].n, L.mf = c0.levels[l].mf;
          L.dinv.reinit(*core.dev, L.n, nb, true);
          L.theta.resize(nb), L.delta.resize(nb);
          for (int b = 0; b < nb; ++b)
            {
              const auto &Lb = clones[b]->get_core().levels[l];
              L.dinv.block(b) = Lb.dinv;
              L.theta[b] = Lb.theta[0], L.delta[b] = Lb.delta[0];
              if (Lb.dinv_mass.size() == 1)
                L.dinv_mass.push_back(Lb.dinv_mass[0]), L.dinv_laplace.push_back(Lb.dinv_laplace[0]);
            }
        }
      const long long n0 = core.levels[0].n;
      core.coarse_inverse.reinit(*core.dev, n0 * n0, nb, true);
      for (int b = 0; b < nb; ++b)
        core.coarse_inverse.block(b) = clones[b]->get_core().coarse_inverse;
      const std::vector<double> m = mass, lp = laplace;
      core.opdesc = [m, lp, nb](unsigned int) { return real_opdesc(nb, m.data(), lp.data()); };
      core.allocate_work_vectors();
    }

    void vmult(Vector &dst, const Vector &src) const { core.vmult(dst, src); }

  private:
    mutable internal::MultigridCore core;
  };
} // namespace spirk_host
