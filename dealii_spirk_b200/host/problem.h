// Host mirror of the reference's namespace HeatEquation (main.cc:2941-3603): Parameters (same JSON
// keys and defaults, 2943-3010) and Problem<dim>::run() split into setup() / step() / finish() so
// that a driver can time individual steps.  Mesh = r-times refined unit hypercube, FE_Q(k),
// homogeneous Dirichlet, manufactured solution sin(2 pi x) sin(2 pi y)[sin(2 pi z)](1+sin pi t)e^{-t/2}.
#pragma once
#include <fstream>
#include <limits>
#include <sstream>

#include "time_integrators.h"

namespace spirk_host
{
  namespace HeatEquation
  {
    struct Parameters
    {
      unsigned int fe_degree     = 4;
      unsigned int n_refinements = 5;

      std::string time_integration_scheme = "ost";
      double      end_time                = 0.5;
      double      time_step_size          = 0.1;

      unsigned int irk_stages                 = 3;
      bool         do_reduce_number_of_vmults = true;

      std::string operator_type             = "MatrixBased";
      std::string block_preconditioner_type = "AMG";

      bool         use_sm       = false;
      bool         do_row_major = true;
      int          padding      = -1;
      unsigned int max_ranks    = 0;

      double outer_tolerance = 1e-8;
      double inner_tolerance = 1e-6;

      bool do_output_paraview = true;

      // extension (not a reference key): keep the reference-literal signs of `ost`
      bool ost_literal_signs = false;

      static bool to_bool(const std::string &s) { return s == "true" || s == "True" || s == "1"; }

      void parse_text(const std::string &text)
      {
        const auto kv = parse_flat_json(text);
        for (const auto &p : kv)
          {
            const std::string &k = p.first, &v = p.second;
            if (k == "FEDegree")
              fe_degree = std::stoul(v);
            else if (k == "NRefinements")
              n_refinements = std::stoul(v);
            else if (k == "TimeIntegrationScheme")
              {
                static const char *allowed[] = {"ost", "irk", "irk_batched", "spirk", "complex_irk", "complex_irk_batched",
                                                "complex_spirk", "complex_spirk_batched"};
                bool ok = false;
                for (auto a : allowed)
                  ok = ok || v == a;
                if (!ok)
                  throw Error("TimeIntegrationScheme: invalid selection '" + v + "'");
                time_integration_scheme = v;
              }
            else if (k == "EndTime")
              end_time = std::stod(v);
            else if (k == "TimeStepSize")
              time_step_size = std::stod(v);
            else if (k == "IRKStages")
              irk_stages = std::stoul(v);
            else if (k == "OuterTolerance")
              outer_tolerance = std::stod(v);
            else if (k == "InnerTolerance")
              inner_tolerance = std::stod(v);
            else if (k == "OperatorType")
              {
                if (v != "MatrixBased" && v != "MatrixFree")
                  throw Error("OperatorType: invalid selection '" + v + "'");
                operator_type = v;
              }
            else if (k == "BlockPreconditionerType")
              {
                if (v != "AMG" && v != "GMG")
                  throw Error("BlockPreconditionerType: invalid selection '" + v + "'");
                block_preconditioner_type = v;
              }
            else if (k == "UseSharedMemory")
              use_sm = to_bool(v);
            else if (k == "DoRowMajor")
              do_row_major = to_bool(v);
            else if (k == "Padding")
              padding = std::stoi(v);
            else if (k == "MaxRanks")
              max_ranks = std::stoul(v);
            else if (k == "DoOutputParaview")
              do_output_paraview = to_bool(v);
            else if (k == "OstLiteralSigns")
              ost_literal_signs = to_bool(v);
            else
              throw Error("parameter file: unknown key '" + k + "'");
          }
      }

      void parse(const std::string file_name)
      {
        std::ifstream file(file_name);
        if (file.fail())
          throw Error("cannot open parameter file " + file_name);
        std::stringstream ss;
        ss << file.rdbuf();
        parse_text(ss.str());
      }
    };

    // the column (space) part of the reference's rectangular process grid (main.cc:3660-3698): this process holds z-slab
    // `rank` of `size` of every partitioned level
    struct ColumnComm
    {
      spirk_comm *comm = nullptr;
      int         rank = 0, size = 1;
    };

    class ProblemBase
    {
    public:
      virtual ~ProblemBase()                          = default;
      virtual void setup()                            = 0;
      virtual bool finished() const                   = 0;
      virtual void step()                             = 0;
      virtual void finish()                           = 0;
      virtual Vector &get_solution()                  = 0;
      virtual const TimeIntegrationSchemes::Interface &integrator() const = 0;

      void run()
      {
        setup();
        while (!finished())
          step();
        finish();
      }

      // per-step records
      std::vector<double> step_time, error_L2, error_Linf, solution_l2, step_seconds;
      double              time            = 0.0;
      unsigned int        timestep_number = 0;
      double              time_step_size  = 0.0;
      long long           n_dofs          = 0; // of the whole mesh
      long long           n_dofs_owned    = 0; // held by this process (a z-slab when the mesh is partitioned)
      long long           first_owned     = 0; // lexicographic index of the first owned DoF
      bool                compute_errors  = true;
    };

    template <int dim>
    class Problem : public ProblemBase
    {
    public:
      Problem(const Parameters &params, Device &device, const TimeIntegrationSchemes::RowComm comm_row, ConvergenceTable &table,
              std::ostream *pcout, const ColumnComm comm_column = ColumnComm())
        : params(params)
        , device(device)
        , comm_row(comm_row)
        , comm_column(comm_column)
        , table(table)
        , pcout(pcout)
      {}

      void setup() override
      {
        if (params.operator_type != "MatrixFree")
          throw Error("OperatorType MatrixBased (Trilinos CSR) is out of scope of this build; use MatrixFree");
        if (params.block_preconditioner_type != "GMG")
          throw Error("BlockPreconditionerType AMG (Trilinos ML) is out of scope of this build; use GMG");
        const std::string &scheme = params.time_integration_scheme;
        const unsigned int r = params.n_refinements, k = params.fe_degree, q = params.irk_stages;

        // Spatial partition (the reference's triangulation on comm_column, main.cc:3027, 3478): level l is split into z-slabs
        // when the plane-streaming cell operator covers it and every slab keeps an even number of cell layers (so that the
        // next coarser level splits at the same planes); coarser levels are held in full by every rank of the column.
        // Levels with fewer than SPIRK_SLAB_MIN_CELLS (default 32) cells per direction are replicated as well: their kernels
        // are latency bound, a halo exchange per operator application would cost more than repeating the work.
        const int          C = comm_column.size;
        const char        *mc = std::getenv("SPIRK_SLAB_MIN_CELLS");
        const unsigned int min_cells = std::max(8, mc ? std::atoi(mc) : 32);
        const auto partitioned = [&](unsigned int l) {
          return C > 1 && dim == 3 && k == 4 && (1u << l) >= std::min(min_cells, 1u << r) && (1u << l) >= 8 && (1u << l) % (2 * C) == 0;
        };
        if (C > 1 && !partitioned(r))
          throw Error("spatial partition: needs 3-D, FEDegree 4 and 2^NRefinements a multiple of twice the number of slabs (>= 8 cells)");
        if (C > 1 && (scheme.rfind("complex", 0) == 0 || scheme == "irk_batched" || comm_row.size * 1u != ((scheme == "spirk") ? q : 1u)))
          throw Error("spatial partition is built for one stage per row rank (spirk with IRKStages row ranks, irk / ost with one)");
        const auto make_level_operator = [&](unsigned int l) {
          if (partitioned(l))
            return std::make_shared<MassLaplaceOperatorMatrixFree<dim>>(device, k, l, comm_column.comm, comm_column.rank, C,
                                                                        /*coarse_replicated=*/l == 0 || !partitioned(l - 1));
          return std::make_shared<MassLaplaceOperatorMatrixFree<dim>>(device, k, l);
        };
        mass_laplace_operator = make_level_operator(r);
        n_dofs_owned          = mass_laplace_operator->m();
        {
          const long long n1 = (long long)k * (1u << r) + 1;
          n_dofs             = (dim == 3) ? n1 * n1 * n1 : n1 * n1;
          first_owned        = (C > 1) ? (long long)k * ((1u << r) * comm_column.rank / C) * n1 * n1 : 0;
        }
        if (pcout)
          *pcout << std::endl
                 << "===========================================" << std::endl
                 << "Number of active cells: " << std::pow((double)(1u << r), dim) << std::endl
                 << "Number of degrees of freedom: " << n_dofs << std::endl
                 << std::endl;
        table.add_value("n_levels", r + 1);
        table.add_value("n_cells", std::pow((double)(1u << r), dim));
        table.add_value("fe_degree", k);
        table.add_value("n_dofs", (double)n_dofs);
        table.add_value("n_stages", q);
        table.add_value("n_procs", comm_row.size * C);
        table.add_value("n_procs_global", comm_row.size * C);
        table.add_value("n_procs_row", comm_row.size);
        table.add_value("n_procs_column", C);

        const bool is_complex = scheme.rfind("complex", 0) == 0;
        if (is_complex)
          complex_mass_laplace_operator =
            std::make_unique<ComplexMassLaplaceOperatorMatrixFree<dim>>(mass_laplace_operator->get_matrix_free());

        // geometric coarsening sequence: levels 0..r, level 0 = one cell (main.cc:3088-3148)
        typename PreconditionerGMG<dim, MassLaplaceOperator>::LevelOperators mg_operators;
        for (unsigned int l = 0; l <= r; ++l)
          {
            auto lop = make_level_operator(l);
            mass_laplace_operator->attach(*lop);
            mg_operators.push_back(lop);
          }
        preconditioner = std::make_unique<PreconditionerGMG<dim, MassLaplaceOperator>>(mg_operators);

        if (scheme == "irk_batched")
          {
            const auto d_vec = load_vector_from_file(q, "D_vec_");
            typename PreconditionerGMG<dim, BatchedMassLaplaceOperator>::LevelOperators bops;
            for (unsigned int l = 0; l <= r; ++l)
              {
                auto bop = std::make_shared<BatchedMassLaplaceOperatorMatrixFree<dim>>(d_vec, mg_operators[l]->get_matrix_free());
                bops.push_back(bop);
                mg_batched_operators.push_back(bop);
              }
            preconditioner_batch = std::make_shared<PreconditionerGMG<dim, BatchedMassLaplaceOperator>>(bops);
          }
        else if (scheme == "complex_irk_batched" || scheme == "complex_spirk_batched")
          {
            typename PreconditionerGMG<dim, ComplexMassLaplaceOperator>::LevelOperators cops;
            for (unsigned int l = 0; l <= r; ++l)
              {
                auto cop = std::make_shared<ComplexMassLaplaceOperatorMatrixFree<dim>>(mg_operators[l]->get_matrix_free());
                complex_mass_laplace_operator->attach(*cop);
                cops.push_back(cop);
              }
            preconditioner_batch = std::make_shared<PreconditionerGMG<dim, ComplexMassLaplaceOperator>>(cops);
          }

        // create_right_hand_side(t) = g(t) * r for the separable forcing (main.cc:3213-3219, 3523-3539)
        mass_laplace_operator->initialize_dof_vector(rhs_spatial);
        SPIRK_CHECK(spirk_problem_rhs_spatial(device.ctx(), &mass_laplace_operator->get_matrix_free().level, rhs_spatial.data()));
        const auto evaluate_rhs_function = [this](const double t, VectorType &tmp) -> void {
          const double pi = 3.14159265358979323846;
          const double g  = (pi * std::cos(pi * t) - 0.5 * (std::sin(pi * t) + 1) + dim * 4.0 * pi * pi * (std::sin(pi * t) + 1)) *
                           std::exp(-0.5 * t);
          if (tmp.size() != rhs_spatial.size())
            tmp.reinit(rhs_spatial, true);
          tmp.equ(g, rhs_spatial);
        };

        namespace TIS = TimeIntegrationSchemes;
        if (scheme == "ost")
          time_integration_scheme =
            std::make_unique<TIS::OneStepTheta>(*mass_laplace_operator, *preconditioner, evaluate_rhs_function, params.ost_literal_signs);
        else if (scheme == "irk" || scheme == "irk_batched")
          time_integration_scheme =
            std::make_unique<TIS::IRK<dim>>(params.outer_tolerance, params.inner_tolerance, q, params.do_reduce_number_of_vmults,
                                            *mass_laplace_operator, *preconditioner, preconditioner_batch, evaluate_rhs_function);
        else if (scheme == "spirk")
          time_integration_scheme = std::make_unique<TIS::IRKStageParallel<dim>>(comm_row, params.outer_tolerance, params.inner_tolerance, q,
                                                                                  params.do_reduce_number_of_vmults, params.use_sm,
                                                                                  *mass_laplace_operator, *preconditioner,
                                                                                  evaluate_rhs_function);
        else if (scheme == "complex_irk" || scheme == "complex_irk_batched")
          time_integration_scheme =
            std::make_unique<TIS::ComplexIRK>(params.outer_tolerance, params.inner_tolerance, q, *mass_laplace_operator,
                                              *complex_mass_laplace_operator, *preconditioner, preconditioner_batch, evaluate_rhs_function);
        else if (scheme == "complex_spirk" || scheme == "complex_spirk_batched")
          time_integration_scheme =
            std::make_unique<TIS::ComplexSPIRK>(comm_row, params.outer_tolerance, params.inner_tolerance, q, *mass_laplace_operator,
                                                *complex_mass_laplace_operator, *preconditioner, preconditioner_batch,
                                                evaluate_rhs_function);
        else
          throw Error("unknown TimeIntegrationScheme");
        time_integration_scheme->pcout = pcout;

        mass_laplace_operator->initialize_dof_vector(solution);
        time = 0.0, timestep_number = 0;
        SPIRK_CHECK(spirk_problem_interpolate_solution(device.ctx(), &level(), solution.data(), 0.0));
        output_results();
        SPIRK_CHECK(spirk_constraints_set_zero(device.ctx(), &level(), 1, solution.data(), solution.size()));

        const double dx = 1.0 / (double)(1u << r); // minimum vertex distance of the uniform mesh
        time_step_size  = (params.time_step_size > 0.0) ? params.time_step_size : std::pow(dx, (k + 1.0) / (2.0 * q - 1.0));
        if (pcout)
          *pcout << std::endl << "Starting time loop with dt=" << time_step_size << std::endl;
        if (!(time_step_size < params.end_time))
          throw Error("time_step_size < end_time required (ExcNotImplemented, ref main.cc:3323)");
      }

      bool finished() const override { return !((params.end_time - time) > (1e-4 * time_step_size)); }

      void step() override
      {
        double time_step_size_truncated = time_step_size;
        if (time + time_step_size > params.end_time)
          {
            const double time_old    = time;
            time                     = params.end_time;
            time_step_size_truncated = time - time_old;
          }
        else
          time += time_step_size;
        if (pcout)
          *pcout << std::endl << "Time step " << timestep_number << " at t=" << time << std::endl;
        ++timestep_number;
        for (auto &bop : mg_batched_operators)
          bop->reinit(time_step_size_truncated);
        device.sync();
        const auto t0 = std::chrono::steady_clock::now();
        time_integration_scheme->solve(solution, timestep_number, time, time_step_size_truncated);
        SPIRK_CHECK(spirk_constraints_set_zero(device.ctx(), &level(), 1, solution.data(), solution.size())); // constraints.distribute
        device.sync();
        step_seconds.push_back(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        output_results();
      }

      void finish() override
      {
        table.add_value("n_t", timestep_number);
        table.add_value("final_t", time);
        table.set_scientific("final_t", true);
        table.add_value("dt", time_step_size);
        table.set_scientific("dt", true);
        table.add_value("error_L2", error_L2.empty() ? 0.0 : error_L2.back());
        table.set_scientific("error_L2", true);
        table.add_value("error_Linf", error_Linf.empty() ? 0.0 : error_Linf.back());
        table.set_scientific("error_Linf", true);
        time_integration_scheme->get_statistics(table, timestep_number > 1 ? timestep_number - 1 : 1);
      }

      Vector &get_solution() override { return solution; }
      const TimeIntegrationSchemes::Interface &integrator() const override { return *time_integration_scheme; }

    private:
      const spirk_level &level() const { return mass_laplace_operator->get_matrix_free().level; }

      void output_results()
      {
        // VTU output (DoOutputParaview) is out of scope; the error norms are the reference's own check
        double l2 = 0, linf = 0;
        if (compute_errors)
          {
            // the cells of this process's slab (they read one ghost plane above), then sum / max over the column
            mass_laplace_operator->get_matrix_free().exchange_ghosts(solution, 0, 1);
            SPIRK_CHECK(spirk_problem_error_norms_partial(device.ctx(), &level(), solution.data(), time, &l2, &linf));
            if (comm_column.size > 1)
              {
                Vector t;
                t.reinit(device, 2, 1, true);
                double v[2] = {l2, linf};
                t.copy_from_host(v);
                SPIRK_CHECK(spirk_comm_allreduce_sum(device.ctx(), comm_column.comm, t.data(), 1));
                SPIRK_CHECK(spirk_comm_allreduce_max(device.ctx(), comm_column.comm, t.data() + 1, 1));
                t.copy_to_host(v);
                l2 = v[0], linf = v[1];
              }
            l2 = std::sqrt(l2);
          }
        step_time.push_back(time);
        error_L2.push_back(l2);
        error_Linf.push_back(linf);
        solution_l2.push_back(compute_errors ? solution.l2_norm() : 0.0);
        if (pcout && compute_errors)
          *pcout << "   Error in the L2/L∞ norm : " << l2 << "/" << linf << std::endl;
      }

      const Parameters                      params;
      Device                               &device;
      const TimeIntegrationSchemes::RowComm comm_row;
      const ColumnComm                      comm_column;
      ConvergenceTable                     &table;
      std::ostream                         *pcout;

      std::shared_ptr<MassLaplaceOperatorMatrixFree<dim>>   mass_laplace_operator;
      std::unique_ptr<ComplexMassLaplaceOperator>           complex_mass_laplace_operator;
      std::unique_ptr<PreconditionerBase<VectorType>>       preconditioner;
      std::shared_ptr<PreconditionerBase<BlockVectorType>>  preconditioner_batch;
      std::vector<std::shared_ptr<const BatchedMassLaplaceOperator>> mg_batched_operators;
      std::unique_ptr<TimeIntegrationSchemes::Interface>    time_integration_scheme;
      VectorType                                            solution, rhs_spatial;
    };
  } // namespace HeatEquation
} // namespace spirk_host
