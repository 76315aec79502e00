// extern "C" surface of the host layer (include/spirk_host.h).
#include <spirk_host.h>

#include <cstring>
#include <sstream>

#include "gmg.h"
#include "problem.h"

using namespace spirk_host;

namespace
{
  thread_local std::string g_host_error;
  template <typename F>
  int guarded(F f)
  {
    try
      {
        f();
        return 0;
      }
    catch (const Error &e)
      {
        g_host_error = e.what();
        return e.status > 0 ? e.status : 1;
      }
    catch (const std::exception &e)
      {
        g_host_error = e.what();
        return 1;
      }
  }
} // namespace

struct spirk_run
{
  std::unique_ptr<Device>                          device;
  spirk_comm                                      *comm = nullptr, *comm_row = nullptr, *comm_col = nullptr;
  HeatEquation::Parameters                         params;
  ConvergenceTable                                 table;
  std::unique_ptr<HeatEquation::ProblemBase>       problem;
  bool                                             verbose = false;
  ~spirk_run()
  {
    problem.reset();
    if (comm_row)
      spirk_comm_destroy(comm_row);
    if (comm_col)
      spirk_comm_destroy(comm_col);
    if (comm)
      spirk_comm_destroy(comm);
    device.reset();
  }
};

extern "C" {

const char *spirk_host_last_error(void) { return g_host_error.c_str(); }
const char *spirk_host_backend(void) { return spirk_backend(); }

int spirk_host_create(const char *json, int is_path, int dim, int device, const char *nccl_id128, int world_rank, int world_size,
                      int verbose, spirk_run **out)
{
  return guarded([&] {
    if (dim != 2 && dim != 3)
      throw Error("dim must be 2 or 3");
    auto run = std::make_unique<spirk_run>();
    if (is_path)
      run->params.parse(json);
    else
      run->params.parse_text(json);
    run->device.reset(new Device(device));
    TimeIntegrationSchemes::RowComm row;
    HeatEquation::ColumnComm        col;
    if (nccl_id128 && world_size > 1)
      {
        SPIRK_CHECK(spirk_comm_create(run->device->ctx(), nccl_id128, world_size, world_rank, &run->comm));
        // The reference's rectangular process grid (main.cc:3660-3698): size_x = stage ranks (IRKStages for spirk, the
        // conjugate pairs for complex_spirk*), the remaining factor of the world size = space ranks (z-slabs of the mesh).
        // With fewer ranks than stages every rank takes several stages and there is one column.
        const std::string &scheme = run->params.time_integration_scheme;
        const int          q      = (int)run->params.irk_stages;
        int                size_x = (scheme == "spirk") ? q : ((scheme == "complex_spirk" || scheme == "complex_spirk_batched") ? (q + 1) / 2 : 1);
        if (world_size < size_x || world_size % size_x != 0)
          {
            if (size_x % world_size != 0)
              throw Error("the number of ranks must be a multiple or a divisor of the number of stage ranks");
            size_x = world_size; // several stages per rank, no spatial partition
          }
        const int size_y = world_size / size_x;
        if (size_y == 1)
          row = TimeIntegrationSchemes::RowComm(run->comm);
        else
          {
            // row-major lex_to_pair (main.cc:281-293): stage index = rank % size_x, slab = rank / size_x
            const int ix = world_rank % size_x, iy = world_rank / size_x;
            SPIRK_CHECK(spirk_comm_split(run->device->ctx(), run->comm, /*color=*/iy, /*key=*/ix, &run->comm_row));
            SPIRK_CHECK(spirk_comm_split(run->device->ctx(), run->comm, /*color=*/ix, /*key=*/iy, &run->comm_col));
            row      = TimeIntegrationSchemes::RowComm(run->comm_row, size_x > 1 ? run->comm : nullptr);
            col.comm = run->comm_col, col.rank = iy, col.size = size_y;
          }
      }
    run->verbose        = verbose && world_rank == 0;
    std::ostream *pcout = run->verbose ? &std::cout : nullptr;
    if (dim == 2)
      run->problem.reset(new HeatEquation::Problem<2>(run->params, *run->device, row, run->table, pcout, col));
    else
      run->problem.reset(new HeatEquation::Problem<3>(run->params, *run->device, row, run->table, pcout, col));
    *out = run.release();
  });
}

int spirk_host_destroy(spirk_run *run)
{
  delete run;
  return 0;
}
int spirk_host_setup(spirk_run *run) { return guarded([&] { run->problem->setup(); }); }
int spirk_host_finished(spirk_run *run, int *finished)
{
  return guarded([&] { *finished = run->problem->finished() ? 1 : 0; });
}
int spirk_host_step(spirk_run *run) { return guarded([&] { run->problem->step(); }); }
int spirk_host_step_host(spirk_run *run, double *host_solution)
{
  return guarded([&] {
    run->problem->get_solution().copy_from_host(host_solution);
    run->problem->step();
    run->problem->get_solution().copy_to_host(host_solution);
  });
}
int spirk_host_finish(spirk_run *run) { return guarded([&] { run->problem->finish(); }); }
int spirk_host_run(spirk_run *run)
{
  return guarded([&] {
    run->problem->run();
    if (run->verbose)
      {
        std::cout << std::endl;
        run->table.write_text(std::cout);
        std::cout << std::endl;
      }
  });
}
int spirk_host_timer_begin(spirk_run *run)
{
  return guarded([&] { SPIRK_CHECK(spirk_ctx_timer_begin(run->device->ctx())); });
}
int spirk_host_timer_end(spirk_run *run, double *ms)
{
  return guarded([&] { SPIRK_CHECK(spirk_ctx_timer_end(run->device->ctx(), ms)); });
}
int spirk_host_set_compute_errors(spirk_run *run, int on)
{
  run->problem->compute_errors = on != 0;
  return 0;
}

int spirk_host_get_scalar(spirk_run *run, const char *key, double *value)
{
  return guarded([&] {
    const std::string k  = key;
    const auto       &p  = *run->problem;
    if (k == "n_dofs")
      *value = (double)p.n_dofs;
    else if (k == "n_dofs_owned")
      *value = (double)p.n_dofs_owned;
    else if (k == "first_owned")
      *value = (double)p.first_owned;
    else if (k == "time")
      *value = p.time;
    else if (k == "timestep_number")
      *value = p.timestep_number;
    else if (k == "dt")
      *value = p.time_step_size;
    else if (k == "n_steps_recorded")
      *value = (double)p.step_seconds.size();
    else if (k == "bytes_allocated")
      *value = (double)run->device->bytes_allocated;
    else if (k == "launch_count")
      *value = (double)spirk_ctx_launch_count(run->device->ctx());
    else
      throw Error("unknown scalar key " + k);
  });
}

int spirk_host_get_array(spirk_run *run, const char *key, double *values, int capacity, int *n)
{
  return guarded([&] {
    const std::string   k = key;
    const auto         &p = *run->problem;
    std::vector<double> v;
    if (k == "step_time")
      v = p.step_time;
    else if (k == "error_L2")
      v = p.error_L2;
    else if (k == "error_Linf")
      v = p.error_Linf;
    else if (k == "solution_l2")
      v = p.solution_l2;
    else if (k == "step_seconds")
      v = p.step_seconds;
    else if (k == "outer_iterations" || k == "inner_iterations")
      {
        if (auto *s = dynamic_cast<const TimeIntegrationSchemes::StatisticsBase *>(&p.integrator()))
          {
            if (k == "outer_iterations")
              for (auto x : s->outer_iterations_per_step)
                v.push_back(x);
            else
              for (auto &row : s->inner_iterations_per_step)
                {
                  double sum = 0;
                  for (auto x : row)
                    sum += x;
                  v.push_back(sum);
                }
          }
        else if (auto *o = dynamic_cast<const TimeIntegrationSchemes::OneStepTheta *>(&p.integrator()))
          v.push_back(o->last_n_iterations);
      }
    else
      throw Error("unknown array key " + k);
    *n = (int)v.size();
    for (int i = 0; i < std::min<int>(capacity, v.size()); ++i)
      values[i] = v[i];
  });
}

int spirk_host_get_solution(spirk_run *run, double *host_solution)
{
  return guarded([&] { run->problem->get_solution().copy_to_host(host_solution); });
}

int spirk_host_table_text(spirk_run *run, char *buffer, int capacity)
{
  return guarded([&] {
    std::ostringstream os;
    run->table.write_text(os);
    std::strncpy(buffer, os.str().c_str(), capacity - 1);
    buffer[capacity - 1] = 0;
  });
}

/* gmg.cc benchmark (one refinement, one mode): fills values[8] = {dim, degree, n_procs, n_cells, n_dofs, L, n_iterations,
 * time per CG iteration [s]} */
int spirk_host_gmg(int dim, int device, int fe_degree, int n_refinements, int mode, int n_components, int n_repetitions,
                   int n_procs, double *values)
{
  return guarded([&] {
    if (dim != 2 && dim != 3)
      throw Error("dim must be 2 or 3");
    if (mode < 0 || mode > 3 || n_components < 1 || n_components > SPIRK_MAX_BLOCKS)
      throw Error("gmg: mode must be 0..3, 1 <= n_components <= SPIRK_MAX_BLOCKS");
    Device                dev(device);
    GMGBenchmark::Result r = (dim == 2) ? GMGBenchmark::test<2>(dev, fe_degree, n_refinements, mode, n_components, n_repetitions, n_procs) :
                                          GMGBenchmark::test<3>(dev, fe_degree, n_refinements, mode, n_components, n_repetitions, n_procs);
    values[0] = r.dim, values[1] = r.degree, values[2] = r.n_procs, values[3] = (double)r.n_cells, values[4] = (double)r.n_dofs;
    values[5] = r.L, values[6] = r.n_iterations, values[7] = r.time;
  });
}
}
