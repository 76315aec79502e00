// Host mirror of the reference's gmg.cc: GMG-preconditioned CG on (M + K) u = 1 in the four batching modes the
// reference measures (gmg.cc:342-382), reported as time per CG iteration with the reference's table columns
// (gmg.cc:293-307).  Isolated measurement of the V-cycle / operator throughput and of "batched vs. one problem
// per process group".
//   mode 0  one component                                   gmg.cc:348-351  test_components<1>(.., 1, ..)
//   mode 1  n components in one vector-valued system         gmg.cc:353-359  test_components<n>(.., n, ..)
//           (FESystem(FE_Q, n): the structured layout stores the components as n blocks; the operator is the
//            batched operator with unit coefficients, the scalar GMG acts on every component)
//   mode 2  n sub-communicators, one component each          gmg.cc:361-375  (here: every rank = one group)
//   mode 3  n components batched (BatchedMassLaplaceOperator + block GMG)   gmg.cc:377-380
#pragma once
#include <chrono>

#include "preconditioner.h"
#include "solvers.h"
#include "tables.h"

namespace spirk_host
{
  namespace GMGBenchmark
  {
    struct Result
    {
      int          dim = 0, degree = 0, n_procs = 1;
      long long    n_cells = 0, n_dofs = 0;
      unsigned int L = 0, n_iterations = 0;
      double       time = 0.0; // seconds per CG iteration (gmg.cc:289-291)
    };

    // the scalar GMG applied to every block of a block vector (mode 1)
    template <typename GMG>
    struct Blockwise
    {
      const GMG &gmg;
      void vmult(Vector &dst, const Vector &src) const
      {
        for (unsigned int b = 0; b < src.n_blocks(); ++b)
          gmg.vmult(dst.block(b), src.block(b));
      }
    };

    template <int dim>
    Result test(Device &device, const unsigned int fe_degree, const unsigned int n_refinements, const int mode,
                const unsigned int n_components, const unsigned int n_repetitions, const int n_procs)
    {
      const unsigned int r = n_refinements, k = fe_degree;
      auto               mass_laplace_operator = std::make_unique<MassLaplaceOperatorMatrixFree<dim>>(device, k, r);
      typename PreconditionerGMG<dim, MassLaplaceOperator>::LevelOperators mg_operators;
      for (unsigned int l = 0; l <= r; ++l)
        {
          auto lop = std::make_shared<MassLaplaceOperatorMatrixFree<dim>>(device, k, l);
          mass_laplace_operator->attach(*lop);
          mg_operators.push_back(lop);
        }
      PreconditionerGMG<dim, MassLaplaceOperator> preconditioner(mg_operators);

      ReductionControl solver_control(1000, 1e-20, 1e-12); // gmg.cc:212
      double           time = 0.0;
      auto             timed = [&](auto &&solve, Vector &dst) {
        dst = 0.0;
        solve(); // warm-up solve (gmg.cc:230-233)
        for (unsigned int counter = 0; counter < n_repetitions; ++counter)
          {
            dst = 0.0;
            device.sync();
            const auto temp = std::chrono::system_clock::now();
            solve();
            device.sync();
            time += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::system_clock::now() - temp).count() / 1e9;
          }
      };

      const bool   single = (mode == 0 || mode == 2);
      const size_t ncomp  = single ? 1 : n_components;
      if (single)
        {
          preconditioner.reinit();
          Vector dst, src;
          mass_laplace_operator->initialize_dof_vector(dst);
          mass_laplace_operator->initialize_dof_vector(src);
          src = 1.0; // gmg.cc:227
          SolverCG cg(solver_control);
          timed([&] { cg.solve(*mass_laplace_operator, dst, src, preconditioner); }, dst);
        }
      else
        {
          const std::vector<double> coeff(n_components, 1.0); // gmg.cc:186-189
          BatchedMassLaplaceOperatorMatrixFree<dim> batched(coeff, mass_laplace_operator->get_matrix_free());
          batched.reinit(1.0);
          Vector dst, src;
          batched.initialize_dof_vector(dst, true);
          batched.initialize_dof_vector(src, true);
          src = 1.0;
          SolverCG cg(solver_control);
          if (mode == 1)
            {
              preconditioner.reinit();
              Blockwise<PreconditionerGMG<dim, MassLaplaceOperator>> blockwise{preconditioner};
              timed([&] { cg.solve(batched, dst, src, blockwise); }, dst);
            }
          else
            {
              typename PreconditionerGMG<dim, BatchedMassLaplaceOperator>::LevelOperators bops;
              for (unsigned int l = 0; l <= r; ++l)
                {
                  auto bop = std::make_shared<BatchedMassLaplaceOperatorMatrixFree<dim>>(coeff, mg_operators[l]->get_matrix_free());
                  bop->reinit(1.0);
                  bops.push_back(bop);
                }
              PreconditionerGMG<dim, BatchedMassLaplaceOperator> preconditioner_batch(bops);
              preconditioner_batch.reinit();
              timed([&] { cg.solve(batched, dst, src, preconditioner_batch); }, dst);
            }
        }

      Result res;
      res.dim = dim, res.degree = (int)k, res.n_procs = n_procs;
      res.n_cells = 1;
      for (int d = 0; d < dim; ++d)
        res.n_cells *= (1LL << r);
      // n_dofs of the (vector-valued) system x number of independent groups (gmg.cc:298-300)
      res.n_dofs       = (long long)mass_laplace_operator->m() * (mode == 1 ? (long long)ncomp : 1) * (mode == 2 ? n_procs : 1);
      res.L            = r + 1;
      res.n_iterations = solver_control.last_step();
      res.time         = time / std::max(1u, solver_control.last_step()) / n_repetitions;
      return res;
    }

    inline void add_to_table(ConvergenceTable &table, const Result &r)
    {
      table.add_value("dim", r.dim);
      table.add_value("degree", r.degree);
      table.add_value("n_procs", r.n_procs);
      table.add_value("n_cells", (double)r.n_cells);
      table.add_value("n_dofs", (double)r.n_dofs);
      table.add_value("L", r.L);
      table.add_value("n_iterations", r.n_iterations);
      table.add_value("time", r.time);
      table.set_scientific("time", true);
    }
  } // namespace GMGBenchmark
} // namespace spirk_host
