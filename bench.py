#!/usr/bin/env python3
"""Benchmark of the dealii-spirk hot path on B200 (contract: see the task statement / DESIGN.md).

A "step" is one fully implicit Runge-Kutta time step of the 3-D Q4 heat equation (rhs assembly,
GMG-preconditioned GMRES on the stage system, solution update).

  N = 1 : BASELINE.json configs[1]: 3-D Q4, IRK q=2, GMG preconditioner, single B200.  The line also carries
          `scaling_reference`: the N>1 workload (spirk q=4) run on this one GPU with all 4 stages batched, the N=1 point of
          the strong-scaling series below.
  N > 1 : stage-parallel SPIRK with a FIXED q = 4 stages - the series BASELINE.json's north_star names ("a SPIRK q=4, 3D Q4
          heat-equation time step on 8xB200 ... per-step time reported at 1/2/4/8 GPUs"): N = 2 two stages per GPU, N = 4
          one stage per GPU (= configs[2]), N = 8 the stage x space grid of 4 stage ranks x 2 z-slabs of the mesh (the
          north-star target configuration).  A strong-scaling series: the outer iteration count depends on q, so only a
          fixed q makes the per-N values comparable.  Fused peer-memory stage mixing, all-reduced Krylov scalars, halo
          exchange of the slabs.  `--stages 8` gives the q = 8 series (8 / 4 / 2 / 1 stages per GPU, profiles/).

metric  : stage-DoFs advanced per second = n_dofs * q / (time per step)   [GDoF*stage/s];
          ms_per_step is the time per SPIRK step; the roofline object reports the dominant kernel
          (the fused Chebyshev cell-operator step) as algorithmic bytes / CUDA-event time against the
          measured HBM copy bandwidth; `vmult_gdofs` is the plain stage-vmult throughput.
value   : device-resident (solution stays in HBM between steps);   e2e: the same step through the
          host-buffer entry point (H2D of u_n from pinned memory + solve + D2H of u_{n+1}).

`--impl reference` times the CPU restatement of the reference (oracle/cpu_abi.cc behind the same
C++ host layer, OpenMP on all host cores) on a bounded sample (smaller refinement) of the workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
TABLES = os.path.join(ROOT, "dealii_spirk_b200", "tables", "butcher_tables.txt")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--refine", type=int, default=6, help="global refinements r: n_dofs = (4*2^r+1)^3")
    ap.add_argument("--degree", type=int, default=4)
    ap.add_argument("--stages", type=int, default=0, help="RK stages q (default: 2 at N=1, 4 otherwise)")
    ap.add_argument("--no-scaling-reference", action="store_true", help="N=1: skip the spirk q=8 run on one GPU")
    ap.add_argument("--scheme", default="", help="irk | spirk | irk_batched | complex_* (default irk at N=1, spirk else)")
    ap.add_argument("--outer-tolerance", type=float, default=1e-8, help="reference default main.cc:2964")
    ap.add_argument("--cpu-refine-min", type=int, default=4, help="smallest refinement of the bounded CPU sample")
    ap.add_argument("--cpu-refine-max", type=int, default=6, help="largest refinement of the bounded CPU sample")
    ap.add_argument("--cpu-budget", type=float, default=0.0,
                    help="seconds of CPU work for the CPU leg (default: 200 for --impl reference, 25 for the cpu_baseline of the GPU arm)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true",
                    help="bracket the timed device-resident steps with cudaProfilerStart/Stop (ncu --profile-from-start off)")
    return ap.parse_args()


def params(scheme, k, r, q, tol, n_steps):
    return {"FEDegree": k, "NRefinements": r, "TimeIntegrationScheme": scheme, "IRKStages": q, "TimeStepSize": 0.1,
            "EndTime": 0.1 * n_steps + 0.05, "OperatorType": "MatrixFree", "BlockPreconditionerType": "GMG",
            "OuterTolerance": tol, "InnerTolerance": 0.0, "DoOutputParaview": False}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.proc = index, [], set(), None
        self.max_mhz = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
        except OSError:
            return
        threading.Thread(target=self._read, daemon=True).start()

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            try:
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except (ValueError, IndexError):
                pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def _native_cpu_libs():
    """The CPU restatement compiled -O3 -march=native ON THIS HOST (directory keyed by the host CPU, so a build made
    on another machine is never executed here); falls back to the portable x86-64-v3 build."""
    import hashlib
    try:
        info = open("/proc/cpuinfo").read()
        key = [l for l in info.splitlines() if l.startswith(("model name", "flags"))][:2]
    except OSError:
        key = []
    tag = hashlib.sha1("\n".join(key).encode()).hexdigest()[:10]
    out = "_build_native_" + tag
    try:
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "native", "NATIVE_OUT=" + out],
                              stdout=subprocess.DEVNULL)
        return os.path.join(ROOT, "oracle", out), "-O3 -march=native"
    except (subprocess.CalledProcessError, OSError):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"])
        return os.path.join(ROOT, "oracle", "_build"), "-O3 -march=x86-64-v3"


def cpu_reference_run(a, q, scheme, n_steps, warmup, budget_s):
    """The restated reference algorithm (oracle/cpu_abi.cc under the same C++ host layer) on the host cores.
    A step is a BOUNDED SAMPLE of the GPU arm's workload: same scheme / q / degree / dt / tolerances at the largest
    refinement r <= a.refine for which warmup + n_steps steps fit into budget_s (a step costs ~8.5x more per
    refinement).  Everything returned describes what was actually run."""
    from dealii_spirk_b200 import capi, hostapi
    libdir, flags = _native_cpu_libs()
    host = hostapi.HostLib(os.path.join(libdir, "libspirk_host_cpu.so"), TABLES)
    cores = os.cpu_count() or 1

    def timed(r, steps, warm):
        t_begin = time.perf_counter()
        with hostapi.Run(host, params(scheme, a.degree, r, q, a.outer_tolerance, steps + warm), dim=3) as run:
            run.set_compute_errors(False)
            run.setup()
            n = run.scalar("n_dofs")
            for _ in range(warm):
                run.step()
            t0 = time.perf_counter()
            for _ in range(steps):
                run.step()
            dt = (time.perf_counter() - t0) / steps
            outer = run.array("outer_iterations")[-steps:]
        return dt, int(n), outer.tolist(), time.perf_counter() - t_begin

    r = min(a.cpu_refine_min, a.refine)
    dt, n, outer, total = timed(r, 1, 1)  # calibration
    spent = total
    steps, warm = n_steps, warmup
    while r < min(a.refine, a.cpu_refine_max) and spent + 8.5 * dt * (steps + warm + 1.5) <= budget_s:
        r, dt = r + 1, 8.5 * dt
    if spent + dt * (steps + warm + 1.5) > budget_s:  # even the smallest sample does not fit: fewer steps, said so
        steps = max(1, min(steps, int((budget_s - spent) / dt) - 2))
        warm = 1
    dt, n, outer, total = timed(r, steps, warm)
    # stage-vmult throughput of the CPU cell loop at the same refinement (the other half of the metric)
    dev = capi.DeviceLib(os.path.join(libdir, "libspirk_cpu.so"))
    lvl = capi.Level(3, a.degree, 2 ** r, 0)
    with capi.Context(dev) as ctx:
        src, dst = ctx.alloc(lvl.n_dofs), ctx.alloc(lvl.n_dofs)
        ctx.call("spirk_vec_set", src, lvl.n_dofs, 0.5)
        op = capi.real_op([16.0], [0.1])
        ctx.call("spirk_op_apply", C.byref(lvl), C.byref(op), dst, src, lvl.n_dofs)
        t0, reps = time.perf_counter(), 3
        for _ in range(reps):
            ctx.call("spirk_op_apply", C.byref(lvl), C.byref(op), dst, src, lvl.n_dofs)
        vm = lvl.n_dofs * reps / (time.perf_counter() - t0) * 1e-9
    return {"value": n * q / dt * 1e-9, "ms_per_step": dt * 1e3, "cores": cores, "n_dofs": n, "refine": r, "steps": steps,
            "warmup": warm, "outer_iterations": outer, "vmult_gdofs": vm, "flags": flags, "scheme": scheme}


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = a.gpus
    q = a.stages or (2 if n_gpus == 1 else 4)
    if q % n_gpus and n_gpus % q:
        raise SystemExit(f"--stages {q} and --gpus {n_gpus}: one must divide the other")
    scheme = a.scheme or ("irk" if n_gpus == 1 else "spirk")
    metric = "implicit RK time step: stage-DoFs advanced per second (n_dofs*q/step time); ms_per_step = time per SPIRK step"
    unit = "GDoF*stage/s"
    config = {"workload": f"3D heat equation Q{a.degree}, {scheme} q={q}, GMG(Chebyshev 5)+GMRES, hypercube r={a.refine}",
              "n_dofs": (a.degree * 2 ** a.refine + 1) ** 3, "stages": q, "refine": a.refine, "degree": a.degree,
              "outer_tolerance": a.outer_tolerance, "dt": 0.1,
              "parallelism": "1 GPU, stages batched" if n_gpus == 1 else (
                  f"stage-parallel: {q // n_gpus} stage(s) per GPU x {n_gpus}" if q % n_gpus == 0 else
                  f"stage x space grid: {q} stage ranks x {n_gpus // q} z-slabs of the mesh"),
              "l2_policy": "working set (>= 30 vectors of n_dofs*8 B) exceeds the 126 MB L2; no explicit flush"}

    if a.impl == "reference":
        if rank != 0:
            return 0
        # one CPU process: the stage-parallel scheme is run as its single-process twin (same algebra, main.cc:815-974)
        cpu_scheme = {"spirk": "irk", "complex_spirk": "complex_irk", "complex_spirk_batched": "complex_irk_batched"}.get(scheme, scheme)
        ref = cpu_reference_run(a, q, cpu_scheme, max(1, a.steps), max(0, a.warmup), a.cpu_budget or 200.0)
        sample = (f"{cpu_scheme} q={q} Q{a.degree} at refinement r={ref['refine']} ({ref['n_dofs']} DoFs) instead of r={a.refine}, "
                  f"{ref['steps']} timed + {ref['warmup']} warm-up steps, {ref['cores']} OpenMP threads, {ref['flags']}; "
                  "restated reference algorithm (oracle/cpu_abi.cc under the same C++ host layer), not deal.II")
        # the line describes what RAN: config / steps / warmup are the sample's, the GPU arm's workload is named beside it
        ref_config = dict(config, workload=f"3D heat equation Q{a.degree}, {cpu_scheme} q={q}, GMG(Chebyshev 5)+GMRES, hypercube "
                          f"r={ref['refine']} (bounded CPU sample of: {config['workload']})",
                          n_dofs=ref["n_dofs"], refine=ref["refine"], parallelism=f"1 CPU process, {ref['cores']} OpenMP threads",
                          l2_policy="n/a (CPU)", sample_of=config["workload"])
        line = {"impl": "reference", "metric": metric, "value": ref["value"], "unit": unit, "n_gpus": n_gpus, "steps": ref["steps"],
                "warmup": ref["warmup"], "ms_per_step": ref["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": ref_config,
                "cpu_baseline": {"value": ref["value"], "unit": unit, "cores": ref["cores"], "kind": "port", "sample": sample},
                "e2e": {"value": ref["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "vmult_gdofs": ref["vmult_gdofs"], "outer_iterations": [int(x) for x in ref["outer_iterations"]],
                "comparable": ref["refine"] == a.refine and cpu_scheme == scheme,
                "note": "restated reference algorithm (cell-loop sum factorisation, deal.II conventions), not deal.II itself; "
                        "a bounded sample: throughput in DoF*stage/s is comparable across refinements, ms_per_step is not"}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    import dealii_spirk_b200 as pkg
    from dealii_spirk_b200 import capi, hostapi

    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = pkg.device_lib()
    host = hostapi.HostLib(pkg.HOST_LIB_PATH, TABLES)
    if host.backend() != "cuda-sm_100a":
        raise SystemExit("the product host library must be linked against the CUDA device library")

    nccl_id = None
    if world > 1:
        buf = C.create_string_buffer(128)
        if rank == 0:
            dev.call("spirk_comm_unique_id", buf)
        t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).cuda()
        dist.broadcast(t, 0)
        nccl_id = bytes(t.cpu().numpy().tobytes())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    total_steps = 2 * (a.steps + a.warmup) + 2
    run = hostapi.Run(host, params(scheme, a.degree, a.refine, q, a.outer_tolerance, total_steps), dim=3, device=local_rank,
                      nccl_id=nccl_id, rank=rank, world=world)
    run.setup()
    n_dofs = run.scalar("n_dofs")
    err0 = run.array("error_L2")
    run.set_compute_errors(False)  # the QGauss(k+2) error evaluation is output, not part of solve()

    sampler = ClockSampler(local_rank)
    # ---- device-resident steps
    for _ in range(a.warmup):
        run.step()
    barrier()
    l0 = run.scalar("launch_count")
    sampler.start()
    # CUDA events on the library's own stream (torch.cuda.Event would only see torch's stream)
    if a.profile:
        torch.cuda.profiler.start()
    run.timer_begin()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        run.step()
    dev_ms = run.timer_end()
    if a.profile:
        torch.cuda.profiler.stop()
    barrier()
    wall = time.perf_counter() - t0
    launches = run.scalar("launch_count") - l0
    step_seconds = run.array("step_seconds")[-a.steps:]
    t_step = max_over_ranks(dev_ms * 1e-3 / a.steps)
    outer = run.array("outer_iterations")[-a.steps:]

    # ---- end-to-end steps through host buffers (pinned staging inside the library call)
    n_owned = run.scalar("n_dofs_owned")  # (a z-slab of the mesh when the ranks outnumber the stages)
    u_pin = torch.empty(int(n_owned), dtype=torch.float64, pin_memory=True)  # pinned host staging
    u_host = u_pin.numpy()
    u_host[:] = run.solution()
    for _ in range(max(1, a.warmup // 2)):
        run.step_host(u_host)
    barrier()
    run.timer_begin()
    for _ in range(a.steps):
        run.step_host(u_host)
    t_e2e = max_over_ranks(run.timer_end() * 1e-3 / a.steps)
    barrier()
    clocks = sampler.stop()
    run.set_compute_errors(True)
    run.step()
    err_final = run.array("error_L2")[-1]

    # ---- dominant kernel, timed live with CUDA events on the library's stream
    roof, vmult_gdofs = None, None
    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        lvl = capi.Level(3, a.degree, 2 ** a.refine, 0)
        N = lvl.n_dofs
        m = q if n_gpus == 1 else max(1, q // n_gpus)
        with capi.Context(dev, local_rank) as ctx:
            x, xo, rhs, dinv, dst = (ctx.alloc(m * N) for _ in range(5))
            ctx.call("spirk_vec_set", x, m * N, 0.5)
            ctx.call("spirk_op_inverse_diagonal", C.byref(lvl), dinv, 16.0, 0.1)
            op = capi.real_op([16.0, 3.16, 2.94, 5.64, 1.0, 2.0, 3.0, 4.0][:m], [0.1])
            f1, _k1 = capi.darr([0.3] * m)
            f2, _k2 = capi.darr([1.1] * m)
            res = {}
            # the smoother's call: D^-1 = the operator's own inverse diagonal, formed on the fly (dinv NULL)
            for name, fn in (("cheb_step", lambda: ctx.call("spirk_op_cheb_step", C.byref(lvl), C.byref(op), dst, x, xo, rhs,
                                                            None, N, f1, f2)),
                             ("vmult", lambda: ctx.call("spirk_op_apply", C.byref(lvl), C.byref(op), dst, x, N))):
                for _ in range(3):
                    fn()
                reps = 20
                ctx.call("spirk_ctx_timer_begin")
                for _ in range(reps):
                    fn()
                res[name] = ctx.scalar_call("spirk_ctx_timer_end") / reps
        # algorithmic bytes per DoF*stage: fused Chebyshev step reads x, x_old, rhs and writes
        # x_new = 32 B (D^-1 is formed on the fly, A x never reaches memory); plain vmult reads src, writes dst = 16 B
        cheb_gbs = 32.0 * m * N / res["cheb_step"] * 1e-6
        vmult_gdofs = m * N / res["vmult"] * 1e-6
        # DRAM bytes per launch of this kernel from the committed `ncu --set full` capture (profiles/), scaled to this launch's
        # algorithmic bytes; NOT measured in this run (bench numbers are never taken under a profiler)
        traffic, traffic_src = None, None
        for name in ("r02_traffic.json", "r01_traffic.json"):
            tpath = os.path.join(ROOT, "profiles", name)
            if os.path.exists(tpath):
                traffic = json.load(open(tpath))["traffic_over_algorithmic"] * 32.0 * m * N
                traffic_src = "profiles/" + name
                break
        roof = {"bound": "hbm", "kernel": "fused Chebyshev step (cell operator + 3-term update), the smoother's kernel",
                "achieved": cheb_gbs, "peak": peak, "unit": "GB/s", "frac": cheb_gbs / peak, "traffic": traffic,
                "traffic_source": traffic_src,
                "traffic_note": "dram read+write bytes per launch from the committed ncu --set full capture (ratio to algorithmic bytes x this launch's algorithmic bytes); not re-measured in this run",
                "peak_source": peak_src, "algorithmic_bytes_per_dof": 32, "launch_ms": res["cheb_step"],
                "vmult": {"achieved": 16.0 * vmult_gdofs, "frac": 16.0 * vmult_gdofs / peak, "launch_ms": res["vmult"],
                          "algorithmic_bytes_per_dof": 16}}
    run.close()

    # ---- N = 1 point of the strong-scaling series: the N > 1 workload (spirk, q = 4) with all stages batched on this GPU
    scaling_ref = None
    if n_gpus == 1 and not a.no_scaling_reference and not a.stages and not a.scheme:
        qs, ns = 4, max(2, min(a.steps, 4))
        with hostapi.Run(host, params("spirk", a.degree, a.refine, qs, a.outer_tolerance, ns + 3), dim=3, device=local_rank) as r8:
            r8.setup()
            r8.set_compute_errors(False)
            for _ in range(2):
                r8.step()
            r8.timer_begin()
            for _ in range(ns):
                r8.step()
            ms8 = r8.timer_end() / ns
            scaling_ref = {"workload": f"3D heat equation Q{a.degree}, spirk q={qs}, hypercube r={a.refine}, all {qs} stages batched on 1 GPU",
                           "stages": qs, "steps": ns, "ms_per_step": ms8, "value": n_dofs * qs / ms8 * 1e-6, "unit": unit,
                           "outer_iterations": [int(x) for x in r8.array("outer_iterations")[-ns:]]}

    cpu = None
    if rank == 0 and n_gpus == 1 and not a.no_cpu_baseline:
        # in a process of its own: the CUDA libraries of this process are loaded RTLD_GLOBAL, a CPU double loaded
        # beside them would bind to their symbols and silently run on the GPU
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2", "--warmup", "1",
                              "--cpu-budget", str(a.cpu_budget or 25.0), "--cpu-refine-min", str(a.cpu_refine_min),
                              "--cpu-refine-max", str(a.cpu_refine_max), "--refine", str(a.refine), "--degree", str(a.degree),
                              "--stages", str(q), "--scheme", scheme, "--outer-tolerance", str(a.outer_tolerance)],
                             capture_output=True, text=True, check=True)
        ref = json.loads(out.stdout.strip().splitlines()[-1])
        cpu = dict(ref["cpu_baseline"], ms_per_step=ref["ms_per_step"], vmult_gdofs=ref["vmult_gdofs"], refine=ref["config"]["refine"],
                   steps=ref["steps"])

    if rank == 0:
        value = n_dofs * q / t_step * 1e-9
        line = {"metric": metric, "value": value, "unit": unit, "n_gpus": n_gpus, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "weak" if n_gpus == 1 else "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config, "clocks": clocks,
                "e2e": {"value": n_dofs * q / t_e2e * 1e-9, "unit": unit, "ms_per_step": t_e2e * 1e3,
                        "h2d_bytes_per_step": int(n_owned * 8), "d2h_bytes_per_step": int(n_owned * 8)},
                "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu, "vmult_gdofs": vmult_gdofs,
                "outer_iterations": [int(x) for x in outer], "step_ms_host_clock": [round(s * 1e3, 3) for s in step_seconds], "wall_ms_per_step": wall / a.steps * 1e3,
                "error_L2_t0": float(err0[0]), "error_L2_final": float(err_final), "scaling_reference": scaling_ref}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
