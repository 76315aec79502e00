// Host mirror of the reference's namespace TimeIntegrationSchemes (main.cc:450-2937):
//   Interface 455-469, OneStepTheta 476-595, IRKBase 663-764, IRK 771-1222,
//   IRKStageParallel 1229-1760, ComplexIRKBase 1767-1879, ComplexIRK 1886-2375,
//   ComplexSPIRK 2382-2934.
// Same class names, constructor meaning, solve()/get_statistics() contract, iteration-count and
// timer bookkeeping, and printed lines.  What changed is how the work is executed:
//  * the stage system matrix (q K-vmults + q M-vmults + q^2 axpys, main.cc:1014-1028) is ONE fused
//    cell pass with a coupled operator descriptor;
//  * the q per-stage V-cycles of the block preconditioner run as one stage-batched V-cycle;
//  * the MPI ring of the stage-parallel classes (main.cc:1443-1534, 2594-2641) is an NCCL all-gather
//    of the stage blocks followed by one mixing kernel; the row all-reduces are NCCL all-reduces.
// A row communicator with P ranks owns q / P consecutive stages (resp. conjugate pairs) per rank, so
// the same classes cover 1 GPU (all stages batched on one device) up to one stage per GPU.
#pragma once
#include <array>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <functional>
#include <iostream>
#include <tuple>

#include "preconditioner.h"
#include "tables.h"

namespace spirk_host
{
  namespace TimeIntegrationSchemes
  {
    using RhsFunction = std::function<void(const double, VectorType &)>;

    // rank / size of the stage ("row") communicator; NULL communicator == single rank
    struct RowComm
    {
      spirk_comm  *comm = nullptr;
      unsigned int rank = 0, size = 1;
      // what the inner products of the stage-coupled vectors are summed over (ReshapedVector, main.cc:237-264): the row
      // communicator, or row x column when the mesh is partitioned as well (nullptr: `comm`)
      spirk_comm *reduce = nullptr;
      spirk_comm *reduction_comm() const { return reduce ? reduce : comm; }
      RowComm()        = default;
      explicit RowComm(spirk_comm *c, spirk_comm *row_x_column = nullptr)
        : comm(c)
        , reduce(row_x_column)
      {
        if (c)
          {
            int r, n;
            SPIRK_CHECK(spirk_comm_rank(c, &r, &n));
            rank = r, size = n;
            if (n == 1)
              comm = nullptr;
          }
      }
      // all[rank*local.size() ...] = local of every rank (replaces the MPI_Sendrecv_replace ring)
      void all_gather(Vector &all, const Vector &local) const
      {
        if (size == 1)
          {
            all = local;
            return;
          }
        SPIRK_CHECK(spirk_comm_allgather(local.ctx(), comm, all.data(), local.data(), local.size()));
      }
      void all_reduce_sum(Vector &v) const
      {
        if (size > 1)
          SPIRK_CHECK(spirk_comm_allreduce_sum(v.ctx(), comm, v.data(), v.size()));
      }
      // {min, avg, max} over the ranks
      std::array<double, 3> min_max_avg(Device &dev, const double x) const
      {
        if (size == 1)
          return {{x, x, x}};
        std::vector<double> mine(size, 0.0), all(size);
        mine[rank] = x;
        Vector t;
        t.reinit(dev, size, 1, true);
        t.copy_from_host(mine.data());
        all_reduce_sum(t);
        t.copy_to_host(all.data());
        double mn = all[0], mx = all[0], s = 0;
        for (const double v : all)
          mn = std::min(mn, v), mx = std::max(mx, v), s += v;
        return {{mn, s / size, mx}};
      }
      double sum(Device &dev, double x) const
      {
        if (size == 1)
          return x;
        Vector t;
        t.reinit(dev, 1, 1, true);
        t.copy_from_host(&x);
        all_reduce_sum(t);
        t.copy_to_host(&x);
        return x;
      }
    };

    // The reference's timers (main.cc:853-969) bracket finished work; kernels here are asynchronous.  The stamps at the
    // beginning and the end of a step always synchronise (`boundary`); the fine-grained ones inside the Krylov loop do so
    // only with SPIRK_SYNC_TIMERS=1 (they would otherwise cost several host-device round trips per GMRES iteration), so by
    // default t_vmult / t_prec_bc / t_prec_solver measure the host's enqueue time and t / t_rhs / t_solver / t_update are exact.
    inline bool sync_timers()
    {
      static const bool v = [] {
        const char *e = std::getenv("SPIRK_SYNC_TIMERS");
        return e && std::atoi(e) != 0;
      }();
      return v;
    }
    inline double now_ns(const Vector &v, const bool boundary = false)
    {
      if (boundary || sync_timers())
        v.device().sync();
      return (double)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::system_clock::now().time_since_epoch()).count();
    }

    // dst_i = [dst_i +] sum_j T(row0 + i, j) src_j for the nrows local rows (cut-off as in the reference)
    inline void mix_rows(Vector &dst, const Vector &src, const FullMatrix &T, unsigned int row0, unsigned int nrows, bool add,
                         double cutoff)
    {
      std::vector<double> rows((size_t)nrows * T.n());
      for (unsigned int i = 0; i < nrows; ++i)
        for (unsigned int j = 0; j < T.n(); ++j)
          rows[(size_t)i * T.n() + j] = T(row0 + i, j);
      SPIRK_CHECK(spirk_mix(dst.ctx(), nrows, T.n(), dst.data(), dst.stride(), src.data(), src.stride(), dst.block_size(),
                            rows.data(), add ? 1 : 0, cutoff));
    }

    class Interface
    {
    public:
      virtual ~Interface() = default;
      virtual void solve(VectorType &solution, const unsigned int timestep_number, const double time, const double time_step) const = 0;
      virtual void get_statistics(ConvergenceTable &table, const double scaling_factor) const = 0;
      std::ostream *pcout = &std::cout; // NULL silences the per-step lines (ConditionalOStream)
    };

    // One-step-theta method (main.cc:476-595).  `literal_signs` keeps the reference's signs
    // (rhs (M + (1-theta) tau K) u, matrix M - theta tau K — indefinite for tau = 0.1);
    // the default is the Crank-Nicolson form that reproduces the manufactured solution.
    class OneStepTheta : public Interface
    {
    public:
      OneStepTheta(const MassLaplaceOperator &system_matrix, const PreconditionerBase<VectorType> &block_preconditioner,
                   const RhsFunction &evaluate_rhs_function, const bool literal_signs = false)
        : theta(0.5)
        , literal_signs(literal_signs)
        , system_matrix(system_matrix)
        , block_preconditioner(block_preconditioner)
        , evaluate_rhs_function(evaluate_rhs_function)
      {}

      void solve(VectorType &solution, const unsigned int, const double time, const double time_step) const override
      {
        VectorType system_rhs, tmp, forcing_terms;
        system_rhs.reinit(solution), tmp.reinit(solution), forcing_terms.reinit(solution);
        const double s = literal_signs ? 1.0 : -1.0;
        system_matrix.vmult(system_rhs, solution, 1.0, s * (1 - theta) * time_step);
        evaluate_rhs_function(time, tmp);
        forcing_terms = tmp;
        forcing_terms *= time_step * theta;
        evaluate_rhs_function(time - time_step, tmp);
        forcing_terms.add(time_step * (1 - theta), tmp);
        system_rhs += forcing_terms;
        system_matrix.reinit(1.0, -s * (theta * time_step));
        if (this->time_step != time_step)
          {
            block_preconditioner.reinit();
            this->time_step = time_step;
          }
        SolverControl solver_control(1000, 1e-8 * system_rhs.l2_norm());
        SolverCG      cg(solver_control);
        cg.solve(system_matrix, solution, system_rhs, block_preconditioner);
        n_iterations += solver_control.last_step();
        last_n_iterations = solver_control.last_step();
        if (pcout)
          *pcout << "   " << solver_control.last_step() << " CG iterations." << std::endl;
      }
      void get_statistics(ConvergenceTable &table, const double scaling_factor) const override
      {
        table.add_value("n_outer_avg", n_iterations / scaling_factor);
      }
      mutable unsigned int last_n_iterations = 0;

    private:
      const double                          theta;
      const bool                            literal_signs;
      const MassLaplaceOperator            &system_matrix;
      const PreconditionerBase<VectorType> &block_preconditioner;
      const RhsFunction                     evaluate_rhs_function;
      mutable double                        time_step    = 0.0;
      mutable double                        n_iterations = 0;
    };

    // common bookkeeping of IRKBase / ComplexIRKBase (main.cc:663-764, 1767-1879)
    class StatisticsBase : public Interface
    {
    public:
      void get_statistics(ConvergenceTable &table, const double scaling_factor = 1.0) const override
      {
        // Utilities::MPI::min_max_avg over the global communicator (main.cc:689-719): all-gather over the stage ranks
        const auto mma = [&](const double value) { return stat_min_max_avg ? stat_min_max_avg(value) : std::array<double, 3>{{value, value, value}}; };
        const auto o = mma(n_outer_iterations / scaling_factor), in = mma(n_inner_iterations / scaling_factor);
        table.add_value("n_outer_min", o[0]), table.add_value("n_outer_avg", o[1]), table.add_value("n_outer_max", o[2]);
        table.add_value("n_inner_min", in[0]), table.add_value("n_inner_avg", in[1]), table.add_value("n_inner_max", in[2]);
        const auto add_time = [&](const std::string label, const double value) {
          table.add_value(label, mma(value)[1] / 1e9);
          table.set_scientific(label, true);
        };
        add_time("t", time_total);
        add_time("t_rhs", time_rhs);
        add_time("t_solver", time_outer_solver);
        add_time("t_update", time_solution_update);
        add_time("t_vmult", time_system_vmult);
        add_time("t_prec_bc", time_preconditioner_bc);
        add_time("t_prec_solver", time_preconditioner_solver);
      }

      // {min, avg, max} of a per-rank value over the ranks; set by the stage-parallel integrators (identity otherwise)
      mutable std::function<std::array<double, 3>(double)> stat_min_max_avg;

      // per-step records (not in the reference's table; used by the parity tests and bench.py)
      mutable std::vector<unsigned int>              outer_iterations_per_step;
      mutable std::vector<std::vector<unsigned int>> inner_iterations_per_step;

    protected:
      void clear_timers() const
      {
        time_total = time_rhs = time_outer_solver = time_solution_update = 0.0;
        time_system_vmult = time_preconditioner_bc = time_preconditioner_solver = 0.0;
        n_outer_iterations = n_inner_iterations = 0;
      }
      mutable double time_total = 0.0, time_rhs = 0.0, time_outer_solver = 0.0, time_solution_update = 0.0;
      mutable double time_system_vmult = 0.0, time_preconditioner_bc = 0.0, time_preconditioner_solver = 0.0;
      mutable double n_outer_iterations = 0, n_inner_iterations = 0;
    };

    class IRKBase : public StatisticsBase
    {
    public:
      IRKBase(const unsigned int n_stages, const bool do_reduce_number_of_vmults, const MassLaplaceOperator &op,
              const PreconditionerBase<VectorType> &block_preconditioner, const RhsFunction &evaluate_rhs_function)
        : n_stages(n_stages)
        , do_reduce_number_of_vmults(do_reduce_number_of_vmults)
        , A_inv(load_matrix_from_file(n_stages, "A_inv"))
        , T(load_matrix_from_file(n_stages, "T"))
        , T_inv(load_matrix_from_file(n_stages, "T_inv"))
        , b_vec(load_vector_from_file(n_stages, "b_vec_"))
        , c_vec(load_vector_from_file(n_stages, "c_vec_"))
        , d_vec(load_vector_from_file(n_stages, "D_vec_"))
        , op(op)
        , block_preconditioner(block_preconditioner)
        , evaluate_rhs_function(evaluate_rhs_function)
      {
        if (!do_reduce_number_of_vmults)
          throw Error("do_reduce_number_of_vmults == false needs MatrixFree vmult_add (ExcNotImplemented, ref operator.h:313-316)");
      }

    protected:
      const unsigned int        n_stages;
      const bool                do_reduce_number_of_vmults;
      const FullMatrix          A_inv, T, T_inv;
      const std::vector<double> b_vec, c_vec, d_vec;
      const MassLaplaceOperator            &op;
      const PreconditionerBase<VectorType> &block_preconditioner;
      const RhsFunction                     evaluate_rhs_function;
    };

    // IRK (main.cc:771-1222) for row.size == 1 and IRKStageParallel (main.cc:1229-1760) otherwise.
    // Rank r of the row communicator owns the stages [r*m, (r+1)*m), m = q / row.size.
    template <int dim>
    class IRKGeneral : public IRKBase
    {
    public:
      using ScalarGMG = PreconditionerGMG<dim, MassLaplaceOperator, Vector>;

      IRKGeneral(const RowComm row, const double outer_tolerance, const double inner_tolerance, const unsigned int n_stages,
                 const bool do_reduce_number_of_vmults, const bool use_sm, const MassLaplaceOperator &op,
                 const PreconditionerBase<VectorType> &block_preconditioner,
                 const std::shared_ptr<PreconditionerBase<BlockVectorType>> &batch_preconditioner,
                 const RhsFunction &evaluate_rhs_function, const char *solver_label)
        : IRKBase(n_stages, do_reduce_number_of_vmults, op, block_preconditioner, evaluate_rhs_function)
        , row(row)
        , batch_preconditioner(batch_preconditioner)
        , n_max_iterations(1000)
        , outer_tolerance(outer_tolerance)
        , inner_tolerance(inner_tolerance)
        , use_sm(use_sm)
        , solver_label(solver_label)
        , times_preconditioner_solver(n_stages, 0.0)
      {
        if (n_stages % row.size != 0)
          throw Error("the number of stages must be a multiple of the number of stage ranks");
        m_local = n_stages / row.size;
        s0      = row.rank * m_local;
      }
      ~IRKGeneral() override
      {
        if (xbuf)
          spirk_comm_xbuf_destroy(xbuf_ctx, xbuf); // the peer-mapped exchange buffer and the IPC mappings of the other ranks
      }

      void get_statistics(ConvergenceTable &table, const double scaling_factor = 1.0) const override
      {
        StatisticsBase::get_statistics(table, scaling_factor);
        for (unsigned int i = 0; i < 10; ++i)
          {
            const std::string label = "t_prec_solver_" + std::to_string(i);
            table.add_value(label, (i < n_stages) ? (times_preconditioner_solver[i] / 1e9) : 0.0);
            table.set_scientific(label, true);
          }
      }

      void solve(VectorType &solution, const unsigned int timestep_number, const double time, const double time_step) const override
      {
        if (this->time_step != time_step)
          preconditioner_ready = false;
        this->time_step = time_step;
        if (!preconditioner_ready)
          setup_preconditioner(solution);
        if (row.size > 1 && !this->stat_min_max_avg)
          {
            Device *d = &solution.device();
            const RowComm r = row;
            this->stat_min_max_avg = [d, r](const double v) { return r.min_max_avg(*d, v); };
          }

        const double t_total = now_ns(solution, true);
        Device      &dev     = solution.device();
        const long long N    = solution.size();

        BlockVectorType system_rhs, system_solution, g;
        VectorType      tmp;
        system_rhs.reinit(dev, N, m_local), system_solution.reinit(dev, N, m_local), g.reinit(dev, N, m_local, true);
        system_rhs.set_reduction_comm(row.reduction_comm()), system_solution.set_reduction_comm(row.reduction_comm());
        tmp.reinit(solution, true);

        // right-hand side g_i = f(t + (c_i - 1) tau) - K u_n, then rhs = (A_inv (x) I) g   (main.cc:867-891, 1343-1349)
        for (unsigned int i = 0; i < m_local; ++i)
          evaluate_rhs_function(time + (c_vec[s0 + i] - 1.0) * time_step, g.block(i));
        op.vmult(tmp, solution, 0.0, -1.0);
        for (unsigned int i = 0; i < m_local; ++i)
          g.block(i).add(1.0, tmp);
        perform_basis_change(system_rhs, g, A_inv, false, 0.0);

        const double t_solver = now_ns(solution, true);
        this->time_rhs += t_solver - t_total;

        ReductionControl solver_control(n_max_iterations, 1e-20, outer_tolerance);
        n_inner.assign(n_stages, 0);
        try
          {
            SolverGMRES  solver(solver_control);
            SystemMatrix sm{*this};
            Preconditioner pc{*this};
            solver.solve(sm, system_solution, system_rhs, pc);
          }
        catch (const SolverControl::NoConvergence &e)
          {
            throw Error(e.what());
          }
        const double t_update = now_ns(solution, true);
        this->time_outer_solver += t_update - t_solver;
        this->n_outer_iterations += solver_control.last_step();
        outer_iterations_per_step.push_back(solver_control.last_step());
        if (row.size > 1) // the counts of the stages owned by the other ranks (a stage is a rank of the reference's row communicator)
          for (unsigned int i = 0; i < n_stages; ++i)
            n_inner[i] = (unsigned int)std::lround(row.sum(dev, (i >= s0 && i < s0 + m_local) ? (double)n_inner[i] : 0.0));
        inner_iterations_per_step.push_back(n_inner);
        double inner_sum = 0;
        for (unsigned int i = 0; i < m_local; ++i)
          inner_sum += n_inner[s0 + i];
        // IRK: the sum over the stages (main.cc:936-943); IRKStageParallel: the count of this rank's stage (main.cc:1398-1401),
        // here the mean over the local stages
        this->n_inner_iterations += (row.size == 1 ? inner_sum : inner_sum / m_local);

        if (pcout)
          {
            *pcout << "   " << solver_control.last_step() << " outer " << solver_label << " iterations and ";
            if (row.size == 1)
              {
                *pcout << n_inner[0];
                if (batch_preconditioner == nullptr)
                  for (unsigned int i = 1; i < n_stages; ++i)
                    *pcout << "+" << n_inner[i];
              }
            else
              {
                // min / avg / max over the row communicator = over the stages (main.cc:1403-1411)
                unsigned int mn = n_inner[0], mx = n_inner[0];
                double       sum = 0;
                for (unsigned int i = 0; i < n_stages; ++i)
                  mn = std::min(mn, n_inner[i]), mx = std::max(mx, n_inner[i]), sum += n_inner[i];
                *pcout << mn << "/" << sum / n_stages << "/" << mx;
              }
            *pcout << " inner CG iterations." << std::endl;
          }

        // u_{n+1} = u_n + tau sum_i b_i k_i (main.cc:959-960, 1416-1426)
        {
          std::vector<double> w(m_local);
          for (unsigned int i = 0; i < m_local; ++i)
            w[i] = time_step * b_vec[s0 + i];
          if (row.size > 1 && row.rank != 0)
            solution = 0.0;
          SPIRK_CHECK(spirk_mix(dev.ctx(), 1, m_local, solution.data(), N, system_solution.data(), system_solution.stride(), N, w.data(), 1, 0.0));
          row.all_reduce_sum(solution);
        }
        const double t_end = now_ns(solution, true);
        this->time_solution_update += t_end - t_update;
        this->time_total += t_end - t_total;
        last_stage_solution = std::move(system_solution);
        if (timestep_number == 1)
          {
            clear_timers(); // the preconditioner is set up in the first time step (main.cc:971-973)
            std::fill(times_preconditioner_solver.begin(), times_preconditioner_solver.end(), 0.0);
          }
      }

      // stage derivatives k_i of the last step (local stages), for the parity tests
      mutable BlockVectorType last_stage_solution;

    private:
      // (T (x) I) with T in {A_inv, T_inv, T}: all-gather of the stage blocks + one mixing kernel
      // (replaces matrix_vector_rol_operation / perform_basis_change, main.cc:1443-1534)
      void perform_basis_change(Vector &dst, const Vector &src, const FullMatrix &Tm, const bool add, const double cutoff) const
      {
        if (row.size == 1)
          {
            mix_rows(dst, src, Tm, 0, n_stages, add, cutoff);
            return;
          }
        if (peer_exchange(src))
          {
            // fused all-gather + mixing: every rank publishes its stage blocks in a peer-mapped exchange buffer and
            // ONE kernel contracts over all stages reading the remote blocks over NVLink (the reference's MPI-3
            // shared-memory variant, main.cc:1506-1533); no gathered copy is written
            const long long n = src.block_size();
            for (unsigned int i = 0; i < m_local; ++i)
              SPIRK_CHECK(spirk_vec_copy(src.ctx(), spirk_comm_xbuf_local(xbuf) + i * n, src.data() + i * src.stride(), n));
            if (xbuf_a2a)
              {
                // all-to-all: every rank contracts its chunk of every stage block for all outputs (full T)
                std::vector<double> full((size_t)Tm.m() * Tm.n());
                for (unsigned int i = 0; i < Tm.m(); ++i)
                  for (unsigned int j = 0; j < Tm.n(); ++j)
                    full[(size_t)i * Tm.n() + j] = Tm(i, j);
                SPIRK_CHECK(spirk_mix_peer_a2a(src.ctx(), row.comm, xbuf, (int)m_local, dst.data(), dst.stride(), n, full.data(),
                                               add ? 1 : 0, cutoff));
                return;
              }
            std::vector<double> rows((size_t)m_local * Tm.n());
            for (unsigned int i = 0; i < m_local; ++i)
              for (unsigned int j = 0; j < Tm.n(); ++j)
                rows[(size_t)i * Tm.n() + j] = Tm(s0 + i, j);
            SPIRK_CHECK(spirk_mix_peer(src.ctx(), row.comm, xbuf, (int)m_local, (int)m_local, dst.data(), dst.stride(), n,
                                       rows.data(), add ? 1 : 0, cutoff));
            return;
          }
        // (contiguous staging: the all-gather needs the stage blocks of a rank back to back)
        gather_send.view_or_pack(src);
        gathered_flat.reinit(src.device(), src.block_size() * n_stages + 1, 1, true);
        SPIRK_CHECK(spirk_comm_allgather(src.ctx(), row.comm, gathered_flat.data(), gather_send.data(), src.block_size() * m_local));
        gathered.view(src.device(), gathered_flat.data(), src.block_size(), n_stages);
        mix_rows(dst, gathered, Tm, s0, m_local, add, cutoff);
      }

      // peer-mapped exchange buffer of the stage group (created on first use; SPIRK_PEER_MIX=0 keeps NCCL all-gather)
      bool peer_exchange(const Vector &src) const
      {
        if (xbuf_state == 0)
          {
            const char *e = std::getenv("SPIRK_PEER_MIX");
            xbuf_state    = 2;
            xbuf_a2a      = !(e && std::atoi(e) == 1); // SPIRK_PEER_MIX=1: gather formulation, default: all-to-all
            if (!(e && std::atoi(e) == 0))
              {
                xbuf_ctx     = src.ctx();
                const int st = spirk_comm_xbuf_create(src.ctx(), row.comm, src.block_size() * m_local, &xbuf);
                // all ranks must agree (a rank without peer access falls back together with the others)
                const double ok = row.sum(src.device(), st == SPIRK_OK ? 0.0 : 1.0);
                if (ok == 0.0)
                  xbuf_state = 1;
                else if (st == SPIRK_OK)
                  {
                    spirk_comm_xbuf_destroy(src.ctx(), xbuf);
                    xbuf = nullptr;
                  }
              }
          }
        return xbuf_state == 1;
      }
      mutable spirk_xbuf *xbuf       = nullptr;
      mutable spirk_ctx  *xbuf_ctx   = nullptr;
      mutable int         xbuf_state = 0; // 0: not tried, 1: peer exchange, 2: NCCL all-gather
      mutable bool        xbuf_a2a   = true;

      struct SystemMatrix
      {
        const IRKGeneral &p;
        // dst_i = tau K v_i + sum_j A_inv(i,j) M v_j   (main.cc:1014-1028, 1580-1592)
        void vmult(Vector &dst, const Vector &src) const
        {
          const double t0 = now_ns(src);
          const auto  &mf = p.op_level();
          // all stages on this rank: one cell pass with the coupled descriptor - where the device library has a
          // plane-streaming kernel for it (pairs); for more stages the mixing + one K v + M w pass below is the fast form
          // (the general cell kernel with atomics is the only one that covers coupled operators of more than two blocks)
          if (p.row.size == 1 && (p.n_stages <= 2 || !mix_then_km()))
            {
              spirk_opdesc d;
              std::memset(&d, 0, sizeof(d));
              d.kind = SPIRK_OP_COUPLED, d.nb = p.n_stages;
              for (unsigned int i = 0; i < p.n_stages; ++i)
                {
                  d.laplace[i] = p.time_step;
                  for (unsigned int j = 0; j < p.n_stages; ++j)
                    d.coupling[i * p.n_stages + j] = p.A_inv(i, j);
                }
              mf.exchange_ghosts_for_operator(src);
              SPIRK_CHECK(spirk_op_apply(src.ctx(), &mf.level, &d, dst.data(), src.data(), src.stride()));
            }
          else
            {
              // M commutes with the stage mixing: w = (A_inv (x) I) v first, then ONE cell pass dst = tau K v + M w
              // (24 B per DoF) instead of the reference's K pass, M pass and mixing with add (main.cc:1582-1591)
              std::vector<double> one(p.m_local, 1.0), tau(p.m_local, p.time_step);
              p.temp.reinit(src, true);
              p.perform_basis_change(p.temp, src, p.A_inv, false, 0.0);
              mf.exchange_ghosts_for_operator(src);
              mf.exchange_ghosts_for_operator(p.temp);
              SPIRK_CHECK(spirk_op_apply_km(src.ctx(), &mf.level, (int)p.m_local, dst.data(), src.data(), p.temp.data(),
                                            src.stride(), tau.data(), one.data()));
            }
          p.time_system_vmult += now_ns(src) - t0;
        }
        static bool mix_then_km()
        {
          static int v = -1;
          if (v < 0)
            {
              const char *e = std::getenv("SPIRK_IRK_COUPLED_PASS"); // 1: the coupled cell pass for any number of stages
              v             = (e && std::atoi(e) == 1) ? 0 : 1;
            }
          return v == 1;
        }
      };

      struct Preconditioner
      {
        const IRKGeneral &p;
        // (T (x) I) diag_i (d_i M + tau K)^-1 (T_inv (x) I)   (main.cc:1095-1173, 1646-1707)
        void vmult(Vector &dst, const Vector &src) const
        {
          const double cut = (p.row.size == 1 || p.use_sm) ? 1e-12 : 0.0;
          const double t0  = now_ns(src);
          p.perform_basis_change(dst, src, p.T_inv, false, cut);
          const double t1 = now_ns(src);
          p.time_preconditioner_bc += t1 - t0;

          p.tmp_vectors.reinit(src, true);
          if (p.batch_preconditioner)
            {
              p.batch_preconditioner->vmult(p.tmp_vectors, dst);
              p.n_inner[0] += 1;
            }
          else if (p.inner_tolerance > 0.0)
            {
              for (unsigned int i = 0; i < p.m_local; ++i)
                {
                  const double     tb0 = now_ns(src);
                  ReductionControl solver_control(100, p.row.size == 1 ? 1e-10 : 1e-20, p.inner_tolerance);
                  SolverCG         solver(solver_control);
                  p.op.reinit(p.d_vec[p.s0 + i], p.time_step);
                  p.tmp_vectors.block(i) = 0.0;
                  solver.solve(p.op, p.tmp_vectors.block(i), dst.block(i), *p.preconditioners[i]);
                  p.n_inner[p.s0 + i] += solver_control.last_step();
                  p.times_preconditioner_solver[p.s0 + i] += now_ns(src) - tb0;
                }
            }
          else
            {
              if (p.stage_batch)
                p.stage_batch->vmult(p.tmp_vectors, dst);
              else
                for (unsigned int i = 0; i < p.m_local; ++i)
                  {
                    p.op.reinit(p.d_vec[p.s0 + i], p.time_step);
                    p.preconditioners[i]->vmult(p.tmp_vectors.block(i), dst.block(i));
                  }
              for (unsigned int i = 0; i < p.m_local; ++i)
                p.n_inner[p.s0 + i] += 1;
            }
          const double t2 = now_ns(src);
          p.time_preconditioner_solver += t2 - t1;
          if (!(p.inner_tolerance > 0.0) && !p.batch_preconditioner)
            for (unsigned int i = 0; i < p.m_local; ++i)
              p.times_preconditioner_solver[p.s0 + i] += (t2 - t1) / p.m_local;

          p.perform_basis_change(dst, p.tmp_vectors, p.T, false, cut);
          p.time_preconditioner_bc += now_ns(src) - t2;
        }
      };

      void setup_preconditioner(const Vector &) const
      {
        preconditioners.clear();
        stage_batch.reset();
        if (batch_preconditioner)
          batch_preconditioner->reinit();
        else
          {
            // one clone per local stage, each set up right after op.reinit(d_i, tau) (main.cc:1083-1089, 1642-1643)
            std::vector<const ScalarGMG *> gmgs;
            std::vector<double>            mass, lap;
            for (unsigned int i = 0; i < m_local; ++i)
              {
                op.reinit(d_vec[s0 + i], time_step);
                preconditioners.push_back(block_preconditioner.clone());
                preconditioners.back()->reinit();
                gmgs.push_back(dynamic_cast<const ScalarGMG *>(preconditioners.back().get()));
                mass.push_back(d_vec[s0 + i]), lap.push_back(time_step);
              }
            bool all_gmg = true;
            for (auto *g : gmgs)
              all_gmg = all_gmg && g != nullptr;
            if (all_gmg && m_local > 1 && !(inner_tolerance > 0.0))
              stage_batch.reset(new StageBatchedGMG<dim>(gmgs, mass, lap));
          }
        preconditioner_ready = true;
      }

      const MatrixFree &op_level() const { return op.get_matrix_free(); }

      const RowComm row;
      const std::shared_ptr<PreconditionerBase<BlockVectorType>> batch_preconditioner;
      const unsigned int n_max_iterations;
      const double       outer_tolerance, inner_tolerance;
      const bool         use_sm;
      const std::string  solver_label;
      unsigned int       m_local = 0, s0 = 0;

      mutable double                    time_step            = 0.0;
      mutable bool                      preconditioner_ready = false;
      mutable std::vector<double>       times_preconditioner_solver;
      mutable std::vector<unsigned int> n_inner;
      mutable std::vector<std::unique_ptr<const PreconditionerBase<VectorType>>> preconditioners;
      mutable std::unique_ptr<StageBatchedGMG<dim>> stage_batch;
      mutable Vector temp, tmp_vectors, gathered, gathered_flat;
      // the local stage blocks back to back (a view of the vector itself when it is contiguous)
      struct Packed
      {
        Vector        owned;
        const double *ptr = nullptr;
        const double *data() const { return ptr; }
        void          view_or_pack(const Vector &v)
        {
          if (v.contiguous())
            {
              ptr = v.data();
              return;
            }
          owned.reinit(v.device(), v.block_size() * v.n_blocks() + 1, 1, true);
          for (unsigned int b = 0; b < v.n_blocks(); ++b)
            SPIRK_CHECK(spirk_vec_copy(v.ctx(), owned.data() + b * v.block_size(), v.data() + b * v.stride(), v.block_size()));
          ptr = owned.data();
        }
      };
      mutable Packed gather_send;
    };

    // the reference's two class names
    template <int dim>
    class IRK : public IRKGeneral<dim>
    {
    public:
      IRK(const double outer_tolerance, const double inner_tolerance, const unsigned int n_stages, const bool do_reduce_number_of_vmults,
          const MassLaplaceOperator &op, const PreconditionerBase<VectorType> &block_preconditioner,
          const std::shared_ptr<PreconditionerBase<BlockVectorType>> &batch_preconditioner, const RhsFunction &evaluate_rhs_function)
        : IRKGeneral<dim>(RowComm(), outer_tolerance, inner_tolerance, n_stages, do_reduce_number_of_vmults, false, op,
                          block_preconditioner, batch_preconditioner, evaluate_rhs_function, "GMRES")
      {}
    };

    template <int dim>
    class IRKStageParallel : public IRKGeneral<dim>
    {
    public:
      IRKStageParallel(const RowComm comm_row, const double outer_tolerance, const double inner_tolerance, const unsigned int n_stages,
                       const bool do_reduce_number_of_vmults, const bool use_sm, const MassLaplaceOperator &op,
                       const PreconditionerBase<VectorType> &block_preconditioner, const RhsFunction &evaluate_rhs_function)
        : IRKGeneral<dim>(comm_row, outer_tolerance, inner_tolerance, n_stages, do_reduce_number_of_vmults, use_sm, op,
                          block_preconditioner, nullptr, evaluate_rhs_function, "GMRES")
      {}
    };

    // ---------------------------------------------------------------------------------------
    // complex variants
    // ---------------------------------------------------------------------------------------
    class ComplexIRKBase : public StatisticsBase
    {
    public:
      ComplexIRKBase(const unsigned int n_stages, const MassLaplaceOperator &op, const PreconditionerBase<VectorType> &block_preconditioner,
                     const RhsFunction &evaluate_rhs_function)
        : n_stages(n_stages)
        , A_inv(load_matrix_from_file(n_stages, "A_inv"))
        , T_re(load_matrix_from_file(n_stages, "T_re"))
        , T_im(load_matrix_from_file(n_stages, "T_im"))
        , T_inv_re(load_matrix_from_file(n_stages, "T_inv_re"))
        , T_inv_im(load_matrix_from_file(n_stages, "T_inv_im"))
        , b_vec(load_vector_from_file(n_stages, "b_vec_"))
        , c_vec(load_vector_from_file(n_stages, "c_vec_"))
        , d_vec_re(load_vector_from_file(n_stages, "D_vec_re_"))
        , d_vec_im(load_vector_from_file(n_stages, "D_vec_im_"))
        , op(op)
        , block_preconditioner(block_preconditioner)
        , evaluate_rhs_function(evaluate_rhs_function)
      {}
      void get_statistics(ConvergenceTable &table, const double scaling_factor = 1.0) const override
      {
        StatisticsBase::get_statistics(table, scaling_factor);
        for (unsigned int i = 0; i < 10; ++i)
          {
            const std::string label = "t_prec_solver_" + std::to_string(i);
            table.add_value(label, 0.0);
            table.set_scientific(label, true);
          }
      }

    protected:
      const unsigned int        n_stages;
      const FullMatrix          A_inv, T_re, T_im, T_inv_re, T_inv_im;
      const std::vector<double> b_vec, c_vec, d_vec_re, d_vec_im;
      const MassLaplaceOperator            &op;
      const PreconditionerBase<VectorType> &block_preconditioner;
      const RhsFunction                     evaluate_rhs_function;
    };

    // ComplexIRK (main.cc:1886-2375) for row.size == 1, ComplexSPIRK (main.cc:2382-2934) otherwise:
    // rank r owns the conjugate pairs [r*mp, (r+1)*mp), mp = ((q+1)/2) / row.size; pair b carries the
    // stages 2b, 2b+1 (only 2b for the unpaired real eigenvalue of odd q).
    class ComplexIRKGeneral : public ComplexIRKBase
    {
    public:
      ComplexIRKGeneral(const RowComm row, const double outer_tolerance, const double inner_tolerance, const unsigned int n_stages,
                        const MassLaplaceOperator &op, const ComplexMassLaplaceOperator &op_complex,
                        const PreconditionerBase<VectorType> &block_preconditioner,
                        const std::shared_ptr<PreconditionerBase<BlockVectorType>> &batch_preconditioner,
                        const RhsFunction &evaluate_rhs_function)
        : ComplexIRKBase(n_stages, op, block_preconditioner, evaluate_rhs_function)
        , row(row)
        , batch_preconditioner(batch_preconditioner)
        , n_max_iterations(1000)
        , outer_tolerance(outer_tolerance)
        , inner_tolerance(inner_tolerance)
        , op_complex(op_complex)
      {
        n_pairs = (n_stages + 1) / 2;
        if (n_pairs % row.size != 0)
          throw Error("the number of conjugate stage pairs must be a multiple of the number of stage ranks");
        mp = n_pairs / row.size;
        p0 = row.rank * mp;
      }

      void solve(VectorType &solution, const unsigned int timestep_number, const double time, const double time_step) const override
      {
        if (this->time_step != time_step)
          {
            preconditioners_batched.clear();
            preconditioners.clear();
          }
        this->time_step = time_step;
        if (preconditioners_batched.empty() && preconditioners.empty())
          for (unsigned int i = 0; i < mp; ++i)
            {
              // literal order of operations: only the scalar operator is re-coefficiented here
              op.reinit(d_vec_re[(p0 + i) * 2] + d_vec_im[(p0 + i) * 2], time_step);
              if (batch_preconditioner)
                {
                  preconditioners_batched.push_back(batch_preconditioner->clone());
                  preconditioners_batched.back()->reinit();
                }
              else
                {
                  preconditioners.push_back(block_preconditioner.clone());
                  preconditioners.back()->reinit();
                }
            }

        const double    t_total = now_ns(solution, true);
        Device         &dev     = solution.device();
        const long long N       = solution.size();
        const unsigned int n_slots = 2 * n_pairs; // stage slots incl. the empty one of an odd q

        // local stage slots: pair (p0+i) -> slots 2i, 2i+1
        BlockVectorType g, system_rhs, system_solution, all;
        VectorType      tmp;
        g.reinit(dev, N, 2 * mp), system_rhs.reinit(dev, N, 2 * mp, true), system_solution.reinit(dev, N, 2 * mp, true);
        tmp.reinit(solution, true);
        op.vmult(tmp, solution, 0.0, -1.0);
        for (unsigned int s = 0; s < 2 * mp; ++s)
          {
            const unsigned int stage = 2 * p0 + s;
            if (stage < n_stages)
              {
                evaluate_rhs_function(time + (c_vec[stage] - 1.0) * time_step, g.block(s));
                g.block(s).add(1.0, tmp);
              }
          }
        // rhs = (A_inv (x) I) g   (ring #0, main.cc:2485-2495)
        gather_slots(all, g, n_slots);
        mix_slots(system_rhs, all, [&](unsigned int i, unsigned int j) { return A_inv(i, j); }, n_slots);

        const double t_solver = now_ns(solution, true);
        this->time_rhs += t_solver - t_total;

        // ---- PreconditionComplex::vmult (main.cc:2129-2226, 2685-2786)
        // apply T_inv: rows 2b of T_inv_re / T_inv_im
        gather_slots(all, system_rhs, n_slots);
        BlockVectorType src_block, dst_block;
        src_block.reinit(dev, N, 2 * mp, true), dst_block.reinit(dev, N, 2 * mp);
        {
          std::vector<double> W((size_t)2 * mp * n_slots, 0.0);
          for (unsigned int i = 0; i < mp; ++i)
            for (unsigned int j = 0; j < n_stages; ++j)
              {
                W[(size_t)(2 * i) * n_slots + j]     = T_inv_re((p0 + i) * 2, j);
                W[(size_t)(2 * i + 1) * n_slots + j] = T_inv_im((p0 + i) * 2, j);
              }
          SPIRK_CHECK(spirk_mix(dev.ctx(), 2 * mp, n_slots, src_block.data(), N, all.data(), N, N, W.data(), 0, 0.0));
        }
        std::vector<std::tuple<unsigned int, unsigned int, unsigned int>> n_iterations(n_pairs, std::make_tuple(0u, 0u, 0u));
        for (unsigned int i = 0; i < mp; ++i)
          {
            const unsigned int pair = p0 + i;
            ReductionControl   solver_control(n_max_iterations, 1e-20, outer_tolerance);
            op_complex.reinit(d_vec_re[pair * 2], d_vec_im[pair * 2], this->time_step);
            Vector src_i, dst_i;
            src_i.view(dev, src_block.block(2 * i).data(), N, 2);
            dst_i.view(dev, dst_block.block(2 * i).data(), N, 2);
            SolverGMRES solver(solver_control);
            try
              {
                if (!preconditioners_batched.empty())
                  {
                    solver.solve(op_complex, dst_i, src_i, *preconditioners_batched[i]);
                    std::get<0>(n_iterations[pair]) += solver_control.last_step();
                    std::get<1>(n_iterations[pair]) += solver_control.last_step() + 1;
                  }
                else
                  {
                    PreconditionPRESB presb(op, *preconditioners[i], inner_tolerance, d_vec_re[pair * 2], d_vec_im[pair * 2],
                                            this->time_step);
                    solver.solve(op_complex, dst_i, src_i, presb);
                    std::get<0>(n_iterations[pair]) += solver_control.last_step();
                    std::get<1>(n_iterations[pair]) += presb.n_iterations.first;
                    std::get<2>(n_iterations[pair]) += presb.n_iterations.second;
                  }
              }
            catch (const SolverControl::NoConvergence &e)
              {
                throw Error(e.what());
              }
            this->n_outer_iterations += std::get<0>(n_iterations[pair]);
            this->n_inner_iterations += std::get<1>(n_iterations[pair]) + std::get<2>(n_iterations[pair]);
          }
        // apply T: k_i = sum_pairs s_j (T_re(i,2j) z_re - T_im(i,2j) z_im), s_j = 2 for pairs, 1 for the real eigenvalue
        gather_slots(all, dst_block, n_slots);
        {
          std::vector<double> W((size_t)2 * mp * n_slots, 0.0);
          for (unsigned int s = 0; s < 2 * mp; ++s)
            {
              const unsigned int stage = 2 * p0 + s;
              if (stage >= n_stages)
                continue;
              for (unsigned int j = 0; j < n_pairs; ++j)
                {
                  const double scaling = (j < (n_stages / 2)) ? 2.0 : 1.0;
                  W[(size_t)s * n_slots + 2 * j]     = scaling * T_re(stage, j * 2);
                  W[(size_t)s * n_slots + 2 * j + 1] = -scaling * T_im(stage, j * 2);
                }
            }
          SPIRK_CHECK(spirk_mix(dev.ctx(), 2 * mp, n_slots, system_solution.data(), N, all.data(), N, N, W.data(), 0, 0.0));
        }
        const double t_update = now_ns(solution, true);
        this->time_outer_solver += t_update - t_solver;

        // iteration bookkeeping / printed line (main.cc:2045-2064, 2540-2559)
        if (row.size > 1)
          for (unsigned int pr = 0; pr < n_pairs; ++pr)
            {
              const bool mine = pr >= p0 && pr < p0 + mp;
              std::get<0>(n_iterations[pr]) = (unsigned int)row.sum(dev, mine ? std::get<0>(n_iterations[pr]) : 0);
              std::get<1>(n_iterations[pr]) = (unsigned int)row.sum(dev, mine ? std::get<1>(n_iterations[pr]) : 0);
              std::get<2>(n_iterations[pr]) = (unsigned int)row.sum(dev, mine ? std::get<2>(n_iterations[pr]) : 0);
            }
        std::vector<unsigned int> outer, inner;
        for (auto &t : n_iterations)
          outer.push_back(std::get<0>(t)), inner.push_back(std::get<1>(t) + std::get<2>(t));
        outer_iterations_per_step.push_back(*std::max_element(outer.begin(), outer.end()));
        pair_outer_iterations_per_step.push_back(outer);
        inner_iterations_per_step.push_back(inner);
        if (pcout)
          {
            *pcout << "   Solved in: ";
            for (unsigned int i = 0; i < n_iterations.size(); ++i)
              {
                *pcout << (i ? ", " : "") << std::get<0>(n_iterations[i]) << " (" << std::get<1>(n_iterations[i]);
                if (!batch_preconditioner)
                  *pcout << "+" << std::get<2>(n_iterations[i]);
                *pcout << ")";
              }
            *pcout << std::endl;
          }

        {
          std::vector<double> w(2 * mp, 0.0);
          for (unsigned int s = 0; s < 2 * mp; ++s)
            if (2 * p0 + s < n_stages)
              w[s] = time_step * b_vec[2 * p0 + s];
          if (row.size > 1 && row.rank != 0)
            solution = 0.0;
          SPIRK_CHECK(spirk_mix(dev.ctx(), 1, 2 * mp, solution.data(), N, system_solution.data(), N, N, w.data(), 1, 0.0));
          row.all_reduce_sum(solution);
        }
        const double t_end = now_ns(solution, true);
        this->time_solution_update += t_end - t_update;
        this->time_total += t_end - t_total;
        if (timestep_number == 1)
          clear_timers();
      }

      mutable std::vector<std::vector<unsigned int>> pair_outer_iterations_per_step;

    private:
      // PRESB preconditioner for [[A, -B], [B, A]], A = l_re M + tau K, B = l_im M (main.cc:2265-2356, 2842-2894)
      class PreconditionPRESB
      {
      public:
        PreconditionPRESB(const MassLaplaceOperator &op, const PreconditionerBase<VectorType> &preconditioner, const double inner_tolerance,
                          const double lambda_re, const double lambda_im, const double tau)
          : op(op)
          , preconditioner(preconditioner)
          , inner_tolerance(inner_tolerance)
          , lambda_re(lambda_re)
          , lambda_im(lambda_im)
          , tau(tau)
          , n_iterations(0, 0)
        {}

        void vmult(BlockVectorType &dst, const BlockVectorType &src) const
        {
          VectorType temp_0;
          temp_0.reinit(src.block(0), true);
          temp_0 = src.block(0);
          temp_0 += src.block(1);
          apply_H_inverse(dst.block(0), temp_0, n_iterations.first);
          op.reinit(lambda_im, 0.0);
          VectorType temp_1;
          temp_1.reinit(src.block(0), true);
          op.vmult(temp_1, dst.block(0));
          temp_1 *= -1.0;
          temp_1 += src.block(1);
          apply_H_inverse(dst.block(1), temp_1, n_iterations.second);
          dst.block(0) -= dst.block(1);
        }
        mutable std::pair<unsigned int, unsigned int> n_iterations;

      private:
        void apply_H_inverse(VectorType &x, const VectorType &rhs, unsigned int &counter) const
        {
          op.reinit(lambda_re + lambda_im, tau);
          if (inner_tolerance == 0.0)
            {
              preconditioner.vmult(x, rhs);
              counter += 1;
            }
          else
            {
              SolverControl reduction_control(100, inner_tolerance);
              SolverCG      solver(reduction_control);
              x = 0.0;
              solver.solve(op, x, rhs, preconditioner);
              counter += reduction_control.last_step();
            }
        }
        const MassLaplaceOperator            &op;
        const PreconditionerBase<VectorType> &preconditioner;
        const double inner_tolerance, lambda_re, lambda_im, tau;
      };

      void gather_slots(Vector &all, const Vector &local, unsigned int n_slots) const
      {
        all.reinit(local.device(), local.block_size(), n_slots, true);
        row.all_gather(all, local);
      }
      // dst(local slot s) = sum_j W(stage(s), j) all_j over real stages j < n_stages
      template <typename F>
      void mix_slots(Vector &dst, const Vector &all, F W, unsigned int n_slots) const
      {
        std::vector<double> M((size_t)2 * mp * n_slots, 0.0);
        for (unsigned int s = 0; s < 2 * mp; ++s)
          if (2 * p0 + s < n_stages)
            for (unsigned int j = 0; j < n_stages; ++j)
              M[(size_t)s * n_slots + j] = W(2 * p0 + s, j);
        SPIRK_CHECK(spirk_mix(dst.ctx(), 2 * mp, n_slots, dst.data(), dst.block_size(), all.data(), all.block_size(), dst.block_size(),
                              M.data(), 0, 0.0));
      }

      const RowComm row;
      const std::shared_ptr<PreconditionerBase<BlockVectorType>> batch_preconditioner;
      const unsigned int n_max_iterations;
      const double       outer_tolerance, inner_tolerance;
      const ComplexMassLaplaceOperator &op_complex;
      unsigned int       n_pairs = 0, mp = 0, p0 = 0;
      mutable double     time_step = 0.0;
      mutable std::vector<std::unique_ptr<const PreconditionerBase<VectorType>>>      preconditioners;
      mutable std::vector<std::unique_ptr<const PreconditionerBase<BlockVectorType>>> preconditioners_batched;
    };

    class ComplexIRK : public ComplexIRKGeneral
    {
    public:
      ComplexIRK(const double outer_tolerance, const double inner_tolerance, const unsigned int n_stages, const MassLaplaceOperator &op,
                 const ComplexMassLaplaceOperator &op_complex, const PreconditionerBase<VectorType> &block_preconditioner,
                 const std::shared_ptr<PreconditionerBase<BlockVectorType>> &batch_preconditioner, const RhsFunction &evaluate_rhs_function)
        : ComplexIRKGeneral(RowComm(), outer_tolerance, inner_tolerance, n_stages, op, op_complex, block_preconditioner,
                            batch_preconditioner, evaluate_rhs_function)
      {}
    };

    class ComplexSPIRK : public ComplexIRKGeneral
    {
    public:
      ComplexSPIRK(const RowComm comm_row, const double outer_tolerance, const double inner_tolerance, const unsigned int n_stages,
                   const MassLaplaceOperator &op, const ComplexMassLaplaceOperator &op_complex,
                   const PreconditionerBase<VectorType> &block_preconditioner,
                   const std::shared_ptr<PreconditionerBase<BlockVectorType>> &batch_preconditioner, const RhsFunction &evaluate_rhs_function)
        : ComplexIRKGeneral(comm_row, outer_tolerance, inner_tolerance, n_stages, op, op_complex, block_preconditioner,
                            batch_preconditioner, evaluate_rhs_function)
      {}
    };
  } // namespace TimeIntegrationSchemes
} // namespace spirk_host
