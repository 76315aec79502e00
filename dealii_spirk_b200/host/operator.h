// Host mirror of the reference's include/operator.h: same class names, method names, argument
// meaning and error behaviour; the arithmetic is the CUDA cell operator behind the C ABI.
//   MassLaplaceOperator                    ref operator.h:15-100
//   MassLaplaceOperatorMatrixFree          ref operator.h:250-460
//   ComplexMassLaplaceOperator[MatrixFree] ref operator.h:463-698
//   BatchedMassLaplaceOperator[MatrixFree] ref operator.h:701-881
// MassLaplaceOperatorMatrixBased (Trilinos CSR, operator.h:104-246) is out of scope.
#pragma once
#include <cstring>

#include "vector.h"

namespace spirk_host
{
  // what MatrixFree<dim,double> is for the reference: the (structured) mesh level
  struct MatrixFree
  {
    Device     *device = nullptr;
    spirk_level level{};
    spirk_comm *column_comm = nullptr; // the space communicator of a z-slab level (the triangulation's communicator, main.cc:3027)
    long long   n_dofs() const { return spirk_level_n_dofs(&level); } // locally owned
    bool        partitioned() const { return ((level.slab >> 8) & 0xff) > 1; }
    int         n1() const { return level.degree * level.n_cells_1d + 1; }
    // update_ghost_values (operator.h:301-306): n_lo planes below / n_hi above the owned range of every block of v
    void exchange_ghosts(double *v, int nb, long long stride, int n_lo, int n_hi) const
    {
      if (partitioned())
        SPIRK_CHECK(spirk_halo_exchange(device->ctx(), column_comm, &level, nb, v, stride, n_lo, n_hi));
    }
    void exchange_ghosts(const Vector &v, int n_lo, int n_hi) const
    {
      exchange_ghosts(const_cast<double *>(v.data()), (int)v.n_blocks(), v.stride(), n_lo, n_hi);
    }
    // what the cell operator reads beyond the owned planes: one cell layer below, one node plane above
    void exchange_ghosts_for_operator(const Vector &v) const { exchange_ghosts(v, level.degree, 1); }
  };

  inline spirk_opdesc real_opdesc(int nb, const double *mass, const double *laplace)
  {
    spirk_opdesc d;
    std::memset(&d, 0, sizeof(d));
    d.kind = SPIRK_OP_REAL, d.nb = nb;
    for (int b = 0; b < nb; ++b)
      d.mass[b] = mass[b], d.laplace[b] = laplace[b];
    return d;
  }

  // Every level operator can describe itself as a C-ABI operator descriptor; the multigrid
  // smoother uses this to fuse the operator into its Chebyshev / residual kernels.
  class LevelOperatorBase
  {
  public:
    virtual ~LevelOperatorBase()                                          = default;
    virtual const MatrixFree &get_matrix_free() const                     = 0;
    virtual spirk_opdesc      descriptor() const                          = 0;
    virtual void              vmult(Vector &dst, const Vector &src) const = 0;
    virtual void              compute_inverse_diagonal(Vector &diagonal) const = 0;
    virtual void              initialize_block_vector(Vector &vec) const  = 0;
    virtual unsigned int      n_blocks() const                            = 0;
  };

  class MassLaplaceOperator : public LevelOperatorBase
  {
  public:
    using Number = double;

    MassLaplaceOperator()
      : mass_matrix_scaling(1.0)
      , laplace_matrix_scaling(1.0)
    {}

    // sets the coefficients on this operator AND on every attached (level) operator
    void reinit(const double mass_matrix_scaling, const double laplace_matrix_scaling) const
    {
      this->mass_matrix_scaling    = mass_matrix_scaling;
      this->laplace_matrix_scaling = laplace_matrix_scaling;
      for (const auto &op : attached_operators)
        op->reinit(this->mass_matrix_scaling, this->laplace_matrix_scaling);
    }

    virtual void initialize_dof_vector(VectorType &vec) const = 0;

    virtual void vmult(VectorType &dst, const VectorType &src, const double mass_matrix_scaling,
                       const double laplace_matrix_scaling) const
    {
      this->reinit(mass_matrix_scaling, laplace_matrix_scaling);
      this->vmult(dst, src);
    }

    virtual unsigned long long m() const = 0;
    virtual Number             el(unsigned int, unsigned int) const = 0;

    void Tvmult(VectorType &dst, const VectorType &src) const { this->vmult(dst, src); }

    void vmult(VectorType &dst, const VectorType &src) const override = 0;

    void vmult_add(VectorType &dst, const VectorType &src, const double mass_matrix_scaling,
                   const double laplace_matrix_scaling) const
    {
      this->reinit(mass_matrix_scaling, laplace_matrix_scaling);
      this->vmult_add(dst, src);
    }
    virtual void vmult_add(VectorType &dst, const VectorType &src) const = 0;

    // dense level matrix (row-major, n x n, host) — replaces get_system_matrix() for the coarse solve
    virtual std::vector<double> get_system_matrix() const = 0;
    virtual bool                supports_sub_communicator() const = 0;

    void attach(const MassLaplaceOperator &other) const { attached_operators.push_back(&other); }

    double get_mass_scaling() const { return mass_matrix_scaling; }
    double get_laplace_scaling() const { return laplace_matrix_scaling; }

  protected:
    mutable double mass_matrix_scaling;
    mutable double laplace_matrix_scaling;
    mutable std::vector<const MassLaplaceOperator *> attached_operators;
  };

  template <int dim, typename Number = double, int n_components = 1>
  class MassLaplaceOperatorMatrixFree : public MassLaplaceOperator
  {
  public:
    // (dof_handler, constraints, quadrature) of the reference collapse to (device, degree, refinement):
    // hypercube, FE_Q(degree), QGauss(degree+1), homogeneous Dirichlet (main.cc:3038-3039, 3400-3411)
    // column_comm / col_rank / col_size: the z-slab of a spatially partitioned level (the reference's triangulation lives
    // on comm_column, main.cc:3027, 3478); coarse_replicated: the next coarser level is held in full by every rank
    MassLaplaceOperatorMatrixFree(Device &device, const unsigned int fe_degree, const unsigned int n_refinements,
                                  spirk_comm *column_comm = nullptr, const int col_rank = 0, const int col_size = 1,
                                  const bool coarse_replicated = false)
    {
      matrix_free.device            = &device;
      matrix_free.level.dim         = dim;
      matrix_free.level.degree      = fe_degree;
      matrix_free.level.n_cells_1d  = 1 << n_refinements;
      matrix_free.level.slab        = (col_size > 1) ? SPIRK_SLAB(col_rank, col_size, coarse_replicated) : 0;
      if (col_size > 1)
        {
          matrix_free.column_comm = column_comm;
          const long long plane   = (long long)matrix_free.n1() * matrix_free.n1();
          device.register_padding(matrix_free.n_dofs(), SPIRK_SLAB_PAD_LO(fe_degree) * plane, SPIRK_SLAB_PAD_HI * plane);
          device.column_comm = column_comm;
        }
    }

    const MatrixFree &get_matrix_free() const override { return matrix_free; }
    unsigned long long m() const override { return matrix_free.n_dofs(); }
    Number el(unsigned int, unsigned int) const override { throw Error("MassLaplaceOperatorMatrixFree::el: ExcNotImplemented"); }

    void initialize_dof_vector(VectorType &vec) const override { vec.reinit(*matrix_free.device, matrix_free.n_dofs(), 1); }
    void initialize_block_vector(Vector &vec) const override { initialize_dof_vector(vec); }
    unsigned int n_blocks() const override { return 1; }

    using MassLaplaceOperator::vmult;

    spirk_opdesc descriptor() const override { return real_opdesc(1, &mass_matrix_scaling, &laplace_matrix_scaling); }

    void vmult(VectorType &dst, const VectorType &src) const override
    {
      const spirk_opdesc d = descriptor();
      matrix_free.exchange_ghosts_for_operator(src);
      SPIRK_CHECK(spirk_op_apply(matrix_free.device->ctx(), &matrix_free.level, &d, dst.data(), src.data(), dst.stride()));
    }

    void vmult_add(VectorType &, const VectorType &) const override
    {
      throw Error("MassLaplaceOperatorMatrixFree::vmult_add: ExcNotImplemented"); // ref operator.h:313-316
    }

    std::vector<double> get_system_matrix() const override
    {
      const long long     n = matrix_free.n_dofs();
      std::vector<double> A((size_t)n * n);
      SPIRK_CHECK(spirk_op_assemble_dense(matrix_free.device->ctx(), &matrix_free.level, mass_matrix_scaling,
                                          laplace_matrix_scaling, A.data()));
      return A;
    }
    bool supports_sub_communicator() const override { return true; }

    void compute_inverse_diagonal(VectorType &diagonal) const override
    {
      this->initialize_dof_vector(diagonal);
      SPIRK_CHECK(spirk_op_inverse_diagonal(matrix_free.device->ctx(), &matrix_free.level, diagonal.data(), mass_matrix_scaling,
                                            laplace_matrix_scaling));
    }

  private:
    MatrixFree matrix_free;
  };

  class ComplexMassLaplaceOperator : public LevelOperatorBase
  {
  public:
    using Number = double;
    ComplexMassLaplaceOperator()
      : lambda_re(1.0)
      , lambda_im(1.0)
      , tau(1.0)
    {}

    virtual void initialize_dof_vector(VectorType &vec, bool block) const = 0;

    void reinit(const double lambda_re, const double lambda_im, const double tau) const
    {
      this->lambda_re = lambda_re;
      this->lambda_im = lambda_im;
      this->tau       = tau;
      for (const auto &op : attached_operators)
        op->reinit(this->lambda_re, this->lambda_im, this->tau);
    }
    void attach(const ComplexMassLaplaceOperator &other) const { attached_operators.push_back(&other); }

    virtual void set_scalar_operator(MassLaplaceOperator &scalar_operator) = 0;
    virtual void Tvmult(BlockVectorType &dst, const BlockVectorType &src) const = 0;
    virtual unsigned long long m() const = 0;
    Number el(unsigned int, unsigned int) const { throw Error("ComplexMassLaplaceOperator::el: ExcNotImplemented"); }

  protected:
    mutable double lambda_re, lambda_im, tau;
    mutable std::vector<const ComplexMassLaplaceOperator *> attached_operators;
  };

  template <int dim, typename Number = double>
  class ComplexMassLaplaceOperatorMatrixFree : public ComplexMassLaplaceOperator
  {
  public:
    ComplexMassLaplaceOperatorMatrixFree(const MatrixFree &matrix_free)
      : matrix_free(matrix_free)
    {}

    void set_scalar_operator(MassLaplaceOperator &op) override { scalar_operator = &op; }
    const MatrixFree &get_matrix_free() const override { return matrix_free; }
    unsigned long long m() const override { return matrix_free.n_dofs() * 2; }
    unsigned int n_blocks() const override { return 2; }

    void initialize_dof_vector(VectorType &vec, bool block) const override
    {
      vec.reinit(*matrix_free.device, matrix_free.n_dofs(), block ? 2 : 1);
    }
    void initialize_block_vector(Vector &vec) const override { initialize_dof_vector(vec, true); }

    // ref operator.h:560-575: inverse diagonal of lambda_re M + tau K copied to both blocks
    void compute_inverse_diagonal(BlockVectorType &diagonal) const override
    {
      initialize_dof_vector(diagonal, true);
      SPIRK_CHECK(spirk_op_inverse_diagonal(matrix_free.device->ctx(), &matrix_free.level, diagonal.block(0).data(), lambda_re, tau));
      diagonal.block(1) = diagonal.block(0);
    }

    // [[lre M + tau K, -lim M], [lim M, lre M + tau K]] in ONE fused cell pass (ref operator.h:616-665)
    spirk_opdesc descriptor() const override
    {
      spirk_opdesc d;
      std::memset(&d, 0, sizeof(d));
      d.kind = SPIRK_OP_COUPLED, d.nb = 2;
      d.laplace[0] = d.laplace[1] = tau;
      d.coupling[0] = lambda_re, d.coupling[1] = -lambda_im;
      d.coupling[2] = lambda_im, d.coupling[3] = lambda_re;
      return d;
    }

    void vmult(BlockVectorType &dst, const BlockVectorType &src) const override
    {
      if (scalar_operator)
        throw Error("ComplexMassLaplaceOperatorMatrixFree: scalar-operator path needs vmult_add (ExcNotImplemented, "
                    "ref operator.h:313-316); the reference never enables it (main.cc:3256-3258)");
      const spirk_opdesc d = descriptor();
      matrix_free.exchange_ghosts_for_operator(src);
      SPIRK_CHECK(spirk_op_apply(matrix_free.device->ctx(), &matrix_free.level, &d, dst.data(), src.data(), dst.stride()));
    }
    void Tvmult(BlockVectorType &, const BlockVectorType &) const override { throw Error("Tvmult: ExcNotImplemented"); }

  private:
    const MatrixFree           matrix_free;
    const MassLaplaceOperator *scalar_operator = nullptr;
  };

  class BatchedMassLaplaceOperator : public LevelOperatorBase
  {
  public:
    using Number = double;
    BatchedMassLaplaceOperator(const std::vector<double> d_vec)
      : tau(1.0)
      , d_vec(d_vec)
    {}
    void reinit(const double tau) const { this->tau = tau; }
    virtual void initialize_dof_vector(VectorType &vec, bool block) const = 0;
    virtual void Tvmult(BlockVectorType &dst, const BlockVectorType &src) const = 0;
    virtual unsigned long long m() const = 0;
    Number el(unsigned int, unsigned int) const { throw Error("BatchedMassLaplaceOperator::el: ExcNotImplemented"); }

  protected:
    mutable double            tau;
    const std::vector<double> d_vec;
  };

  template <int dim, typename Number = double>
  class BatchedMassLaplaceOperatorMatrixFree : public BatchedMassLaplaceOperator
  {
  public:
    BatchedMassLaplaceOperatorMatrixFree(const std::vector<double> d_vec, const MatrixFree &matrix_free)
      : BatchedMassLaplaceOperator(d_vec)
      , matrix_free(matrix_free)
    {}
    const MatrixFree &get_matrix_free() const override { return matrix_free; }
    unsigned long long m() const override { return matrix_free.n_dofs() * d_vec.size(); }
    unsigned int n_blocks() const override { return d_vec.size(); }

    void initialize_dof_vector(VectorType &vec, bool block) const override
    {
      vec.reinit(*matrix_free.device, matrix_free.n_dofs(), block ? (int)d_vec.size() : 1);
    }
    void initialize_block_vector(Vector &vec) const override { initialize_dof_vector(vec, true); }

    void compute_inverse_diagonal(BlockVectorType &diagonal) const override
    {
      initialize_dof_vector(diagonal, true);
      for (unsigned int b = 0; b < d_vec.size(); ++b)
        SPIRK_CHECK(spirk_op_inverse_diagonal(matrix_free.device->ctx(), &matrix_free.level, diagonal.block(b).data(), d_vec[b], tau));
    }

    spirk_opdesc descriptor() const override
    {
      std::vector<double> lap(d_vec.size(), tau);
      return real_opdesc((int)d_vec.size(), d_vec.data(), lap.data());
    }

    // all stages in one cell pass (ref operator.h:841-880)
    void vmult(BlockVectorType &dst, const BlockVectorType &src) const override
    {
      const spirk_opdesc d = descriptor();
      matrix_free.exchange_ghosts_for_operator(src);
      SPIRK_CHECK(spirk_op_apply(matrix_free.device->ctx(), &matrix_free.level, &d, dst.data(), src.data(), dst.stride()));
    }
    void Tvmult(BlockVectorType &, const BlockVectorType &) const override { throw Error("Tvmult: ExcNotImplemented"); }

  private:
    const MatrixFree matrix_free;
  };
} // namespace spirk_host
