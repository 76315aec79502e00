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
  spirk_comm                                      *comm = nullptr;
  HeatEquation::Parameters                         params;
  ConvergenceTable                                 table;
  std::unique_ptr<HeatEquation::ProblemBase>       problem;
  bool                                             verbose = false;
  ~spirk_run()
  {
    problem.reset();
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
    if (nccl_id128 && world_size > 1)
      {
        SPIRK_CHECK(spirk_comm_create(run->device->ctx(), nccl_id128, world_size, world_rank, &run->comm));
        row = TimeIntegrationSchemes::RowComm(run->comm);
      }
    run->verbose        = verbose && world_rank == 0;
    std::ostream *pcout = run->verbose ? &std::cout : nullptr;
    if (dim == 2)
      run->problem.reset(new HeatEquation::Problem<2>(run->params, *run->device, row, run->table, pcout));
    else
      run->problem.reset(new HeatEquation::Problem<3>(run->params, *run->device, row, run->table, pcout));
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
