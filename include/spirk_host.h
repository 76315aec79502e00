/*
 * spirk_host.h — C entry points of the C++ host layer (dealii_spirk_b200/host/), which mirrors the
 * reference's operator / preconditioner / time-integrator classes (include/operator.h,
 * include/preconditioner.h, main.cc) above the device C ABI (spirk_b200.h).  These entry points
 * are what a driver in another language binds (bench.py and the tests use ctypes); a C++ caller
 * uses the classes in dealii_spirk_b200/host/*.h directly.
 *
 * One spirk_run == one HeatEquation::Problem<dim> (main.cc:3014-3603) configured from a JSON
 * parameter file with the reference's key set (main.cc:2970-3009).
 */
#ifndef SPIRK_HOST_H
#define SPIRK_HOST_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct spirk_run spirk_run;

const char *spirk_host_last_error(void);
/* name of the device library this host library is linked against ("cuda-sm_100a") */
const char *spirk_host_backend(void);

/* json: parameter text (is_path == 0) or file name (is_path != 0).  dim: 2 or 3 (the reference's
 * compile-time IRK_DIMENSION).  nccl_id128: NULL for a single process, else the 128-byte id from
 * spirk_comm_unique_id shared by all world_size ranks; ranks form the stage ("row") communicator.
 * verbose: print the reference's per-step lines on rank 0. */
int spirk_host_create(const char *json, int is_path, int dim, int device, const char *nccl_id128, int world_rank,
                      int world_size, int verbose, spirk_run **out);
int spirk_host_destroy(spirk_run *run);
/* Problem::run() in three pieces (main.cc:3035-3371) */
int spirk_host_setup(spirk_run *run);
int spirk_host_finished(spirk_run *run, int *finished);
int spirk_host_step(spirk_run *run); /* solution stays resident on the device */
/* end-to-end variant of one step through host buffers: copy host_solution (n_dofs doubles) to the
 * device, advance one step (TimeIntegrationSchemes::Interface::solve), copy the new solution back */
int spirk_host_step_host(spirk_run *run, double *host_solution);
int spirk_host_finish(spirk_run *run);
/* whole run: setup + time loop + statistics table printed to stdout (the reference's main()) */
int spirk_host_run(spirk_run *run);

/* CUDA events on the run's stream bracketing a region: end returns elapsed device milliseconds */
int spirk_host_timer_begin(spirk_run *run);
int spirk_host_timer_end(spirk_run *run, double *ms);

/* switch the per-step error evaluation (QGauss(k+2) integrate_difference) off / on */
int spirk_host_set_compute_errors(spirk_run *run, int on);
/* scalar queries: "n_dofs", "time", "timestep_number", "dt", "n_steps_recorded", "bytes_allocated",
 * "launch_count"; per-step arrays: "step_time", "error_L2", "error_Linf", "solution_l2",
 * "step_seconds", "outer_iterations", "inner_iterations" */
int spirk_host_get_scalar(spirk_run *run, const char *key, double *value);
int spirk_host_get_array(spirk_run *run, const char *key, double *values, int capacity, int *n);
int spirk_host_get_solution(spirk_run *run, double *host_solution);
/* the ConvergenceTable as text (columns of main.cc:689-719, 3360-3368, 3387-3398) */
int spirk_host_table_text(spirk_run *run, char *buffer, int capacity);

/* The reference's gmg.cc benchmark for one refinement level and one of its four modes (gmg.cc:342-382):
 * 0 = one component, 1 = n_components in one vector-valued system, 2 = one component per process group
 * (n_procs groups), 3 = n_components batched.  GMG-preconditioned CG on (M + K) u = 1 to a 1e-12 reduction,
 * one warm-up solve + n_repetitions timed solves.  values[8] = {dim, degree, n_procs, n_cells, n_dofs, L,
 * n_iterations, seconds per CG iteration} (the table columns of gmg.cc:293-307). */
int spirk_host_gmg(int dim, int device, int fe_degree, int n_refinements, int mode, int n_components, int n_repetitions,
                   int n_procs, double *values);

#ifdef __cplusplus
}
#endif
#endif
