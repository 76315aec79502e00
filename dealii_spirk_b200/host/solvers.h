// Krylov solvers with deal.II's semantics (the reference instantiates deal.II's SolverCG /
// SolverGMRES: main.cc:527-536, 920-924, 1131-1138, 1379-1383, 1676-1683, 2170-2175, 2200-2202,
// 2303-2306, 2726-2731; conventions restated in SURVEY Appendix A5/A6).  The loops run on the
// host; every vector operation is a CUDA kernel; each inner product is one device reduction.
#pragma once
#include <algorithm>
#include <cmath>
#include <functional>
#include <limits>

#include "vector.h"

namespace spirk_host
{
  class SolverControl
  {
  public:
    enum State
    {
      iterate,
      success,
      failure
    };
    class NoConvergence : public Error
    {
    public:
      NoConvergence(unsigned int last_step, double last_residual)
        : Error("Iterative method reported convergence failure in step " + std::to_string(last_step) +
                ". The residual in the last step was " + std::to_string(last_residual) + ".")
        , last_step(last_step)
        , last_residual(last_residual)
      {}
      unsigned int last_step;
      double       last_residual;
    };

    SolverControl(const unsigned int n = 100, const double tol = 1.e-10)
      : maxsteps(n)
      , tol(tol)
    {}
    virtual ~SolverControl() = default;

    virtual State check(const unsigned int step, const double check_value)
    {
      lstep = step, lvalue = check_value;
      if (step == 0)
        initial_val = check_value;
      if (check_value <= tol)
        return lcheck = success;
      if (step >= maxsteps || std::isnan(check_value))
        return lcheck = failure;
      return lcheck = iterate;
    }
    unsigned int last_step() const { return lstep; }
    double       last_value() const { return lvalue; }
    double       initial_value() const { return initial_val; }
    State        last_check() const { return lcheck; }

  protected:
    unsigned int maxsteps;
    double       tol;
    unsigned int lstep       = 0;
    double       lvalue      = 0;
    double       initial_val = 0;
    State        lcheck      = iterate;
  };

  class ReductionControl : public SolverControl
  {
  public:
    ReductionControl(const unsigned int n = 100, const double tolerance = 1.e-10, const double reduce = 1.e-2)
      : SolverControl(n, tolerance)
      , reduce(reduce)
    {}
    State check(const unsigned int step, const double check_value) override
    {
      if (step == 0)
        {
          initial_val = check_value;
          reduced_tol = reduce * check_value;
        }
      if (check_value <= reduced_tol)
        {
          lstep = step, lvalue = check_value;
          return lcheck = success;
        }
      return SolverControl::check(step, check_value);
    }

  private:
    double reduce;
    double reduced_tol = 0;
  };

  // deal.II SolverCG::solve (A5).  MatrixType / PreconditionerType need vmult(dst, src).
  class SolverCG
  {
  public:
    explicit SolverCG(SolverControl &cn)
      : control(cn)
    {}

    // Lanczos coefficients of the CG run (what connect_eigenvalues_slot delivers in deal.II)
    std::vector<double> lanczos_diagonal, lanczos_offdiagonal;
    bool                collect_lanczos = false;

    template <typename MatrixType, typename PreconditionerType>
    void solve(const MatrixType &A, Vector &x, const Vector &b, const PreconditionerType &preconditioner)
    {
      Vector g, d, h;
      g.reinit(x, true), d.reinit(x, true), h.reinit(x, true);
      int    it  = 0;
      double res = -std::numeric_limits<double>::max();
      double eigen_beta_alpha = 0, gh, beta = 0, alpha, old_alpha = 0;
      lanczos_diagonal.clear(), lanczos_offdiagonal.clear();

      if (!x.all_zero())
        {
          A.vmult(g, x);
          g.add(-1., b);
        }
      else
        g.equ(-1., b);
      res = g.l2_norm();

      SolverControl::State conv = control.check(0, res);
      if (conv != SolverControl::iterate)
        {
          if (conv == SolverControl::failure)
            throw SolverControl::NoConvergence(0, res);
          return;
        }
      preconditioner.vmult(h, g);
      d.equ(-1., h);
      gh = g * h;

      while (conv == SolverControl::iterate)
        {
          it++;
          A.vmult(h, d);
          alpha = d * h;
          alpha = gh / alpha;
          x.add(alpha, d);
          res = std::sqrt(std::abs(g.add_and_dot(alpha, h, g)));
          if (it > 1 && collect_lanczos)
            {
              lanczos_diagonal.push_back(1. / old_alpha + eigen_beta_alpha);
              eigen_beta_alpha = beta / old_alpha;
              lanczos_offdiagonal.push_back(std::sqrt(beta) / old_alpha);
            }
          conv = control.check(it, res);
          if (conv != SolverControl::iterate)
            break;
          preconditioner.vmult(h, g);
          beta = gh;
          gh   = g * h;
          beta = gh / beta;
          d.sadd(beta, -1., h);
          old_alpha = alpha;
        }
      if (conv != SolverControl::success)
        throw SolverControl::NoConvergence(control.last_step(), control.last_value());
    }

  private:
    SolverControl &control;
  };

  // deal.II SolverGMRES::solve with the default AdditionalData (A6): 30 temporary vectors
  // (restart length 28), left preconditioning, stopping test on the preconditioned residual,
  // modified Gram-Schmidt (one fused device sweep per new Krylov vector) with the Kelley
  // re-orthogonalisation test every fifth iteration.
  class SolverGMRES
  {
  public:
    explicit SolverGMRES(SolverControl &cn, const unsigned int max_n_tmp_vectors = 30)
      : control(cn)
      , n_tmp(max_n_tmp_vectors)
    {}

    template <typename MatrixType, typename PreconditionerType>
    void solve(const MatrixType &A, Vector &x, const Vector &b, const PreconditionerType &preconditioner)
    {
      std::vector<std::unique_ptr<Vector>> tmp(n_tmp);
      auto vec = [&](unsigned int i) -> Vector & {
        if (!tmp[i])
          {
            tmp[i].reset(new Vector());
            tmp[i]->reinit(x, true);
          }
        return *tmp[i];
      };
      unsigned int         accumulated_iterations = 0;
      std::vector<double>  gamma(n_tmp), ci(n_tmp - 1), si(n_tmp - 1), h(n_tmp - 1);
      std::vector<double>  H((size_t)n_tmp * (n_tmp - 1), 0.0); // H(i,j) = H[i*(n_tmp-1)+j]
      unsigned int         dim             = 0;
      SolverControl::State iteration_state = SolverControl::iterate;
      bool                 re_orthogonalize = false;
      Vector              &v = vec(0);
      Vector              &p = vec(n_tmp - 1);

      do
        {
          std::fill(h.begin(), h.end(), 0.0);
          A.vmult(p, x);
          p.sadd(-1., 1., b);
          preconditioner.vmult(v, p);
          double rho      = v.l2_norm();
          iteration_state = control.check(accumulated_iterations, rho);
          if (iteration_state != SolverControl::iterate)
            break;
          gamma[0] = rho;
          v *= 1. / rho;

          for (unsigned int inner_iteration = 0; (inner_iteration < n_tmp - 2) && (iteration_state == SolverControl::iterate);
               ++inner_iteration)
            {
              ++accumulated_iterations;
              Vector &vv = vec(inner_iteration + 1);
              A.vmult(p, vec(inner_iteration));
              preconditioner.vmult(vv, p);
              dim = inner_iteration + 1;

              const double s = modified_gram_schmidt(tmp, dim, accumulated_iterations, vv, h, re_orthogonalize);
              h[inner_iteration + 1] = s;
              if (s != 0)
                vv *= 1. / s;

              givens_rotation(h, gamma, ci, si, inner_iteration);
              for (unsigned int i = 0; i < dim; ++i)
                H[(size_t)i * (n_tmp - 1) + inner_iteration] = h[i];
              rho             = std::fabs(gamma[dim]);
              iteration_state = control.check(accumulated_iterations, rho);
            }
          // solve the triangular system H1 y = gamma (FullMatrix::backward)
          std::vector<double> y(dim);
          for (int i = (int)dim - 1; i >= 0; --i)
            {
              double s = gamma[i];
              for (unsigned int j = i + 1; j < dim; ++j)
                s -= H[(size_t)i * (n_tmp - 1) + j] * y[j];
              y[i] = s / H[(size_t)i * (n_tmp - 1) + i];
            }
          for (unsigned int i = 0; i < dim; ++i)
            x.add(y[i], vec(i));
        }
      while (iteration_state == SolverControl::iterate);

      if (iteration_state != SolverControl::success)
        throw SolverControl::NoConvergence(control.last_step(), control.last_value());
    }

  private:
    static void givens_rotation(std::vector<double> &h, std::vector<double> &b, std::vector<double> &ci, std::vector<double> &si,
                                int col)
    {
      for (int i = 0; i < col; i++)
        {
          const double s = si[i], c = ci[i], dummy = h[i];
          h[i]     = c * dummy + s * h[i + 1];
          h[i + 1] = -s * dummy + c * h[i + 1];
        }
      const double r = 1. / std::sqrt(h[col] * h[col] + h[col + 1] * h[col + 1]);
      si[col]        = h[col + 1] * r;
      ci[col]        = h[col] * r;
      h[col]         = ci[col] * h[col] + si[col] * h[col + 1];
      b[col + 1]     = -si[col] * b[col];
      b[col] *= ci[col];
    }

    static double mgs_sweep(const std::vector<std::unique_ptr<Vector>> &basis, unsigned int dim, Vector &vv, double *h)
    {
      std::vector<const double *> ptrs(dim);
      for (unsigned int i = 0; i < dim; ++i)
        ptrs[i] = basis[i]->data();
      double norm = 0;
      if (vv.reduction_comm())
        spirk_ctx_set_reduction_comm(vv.ctx(), vv.reduction_comm());
      const int st = vv.contiguous() ?
                       spirk_gmres_mgs(vv.ctx(), vv.data(), ptrs.data(), (int)dim, vv.size(), h, &norm) :
                       spirk_gmres_mgs_strided(vv.ctx(), vv.data(), ptrs.data(), (int)dim, vv.block_size(), (int)vv.n_blocks(), vv.stride(),
                                               h, &norm);
      if (vv.reduction_comm())
        spirk_ctx_set_reduction_comm(vv.ctx(), nullptr);
      check(st, "spirk_gmres_mgs");
      return norm;
    }

    static double modified_gram_schmidt(const std::vector<std::unique_ptr<Vector>> &orthogonal_vectors, const unsigned int dim,
                                        const unsigned int accumulated_iterations, Vector &vv, std::vector<double> &h,
                                        bool &re_orthogonalize)
    {
      double     norm_vv_start            = 0;
      const bool consider_reorthogonalize = (re_orthogonalize == false) && (accumulated_iterations % 5 == 0);
      if (consider_reorthogonalize)
        norm_vv_start = vv.l2_norm();
      double norm_vv = mgs_sweep(orthogonal_vectors, dim, vv, h.data());
      if (consider_reorthogonalize)
        {
          if (norm_vv > 10. * norm_vv_start * std::sqrt(std::numeric_limits<double>::epsilon()))
            return norm_vv;
          re_orthogonalize = true;
        }
      if (re_orthogonalize == true)
        {
          std::vector<double> h2(dim);
          norm_vv = mgs_sweep(orthogonal_vectors, dim, vv, h2.data());
          for (unsigned int i = 0; i < dim; ++i)
            h[i] += h2[i];
        }
      return norm_vv;
    }

    SolverControl &control;
    unsigned int   n_tmp;
  };
} // namespace spirk_host
