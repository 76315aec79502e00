#!/usr/bin/env python3
"""Stage-vmult micro-benchmark: GDoF/s of the cell operator and of the fused
Chebyshev step on 3-D Q4, CUDA events on the library's stream, inputs larger than L2 at r>=6."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dealii_spirk_b200 as pkg  # noqa: E402
from dealii_spirk_b200 import capi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=3)
    ap.add_argument("--degree", type=int, default=4)
    ap.add_argument("--refine", type=int, nargs="+", default=[5, 6])
    ap.add_argument("--nb", type=int, default=1)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--variants", type=int, nargs="+", default=[0, 1])
    ap.add_argument("--opt", nargs="*", default=[], help="spirk_ctx_set_option knobs, name=value")
    ap.add_argument("--lib", default="", help="an experiment build of the device library (build.build_variant)")
    ap.add_argument("--kernels", nargs="*", default=[], help="subset of apply cheb_step cheb_step_own_dinv residual")
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    dev = capi.DeviceLib(a.lib) if a.lib else pkg.device_lib()
    peaks = json.load(open(os.path.join(os.path.dirname(pkg.HERE), "MEASURED_PEAKS.json"))) \
        if os.path.exists(os.path.join(os.path.dirname(pkg.HERE), "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    for r in a.refine:
        lvl = capi.Level(a.dim, a.degree, 2 ** r, 0)
        N = lvl.n_dofs
        with capi.Context(dev) as ctx:
            i = np.arange(a.nb * N, dtype=np.float64) + 1.0
            v = np.sin(12.9898 * i) * 43758.5453
            src = ctx.upload(2.0 * (v - np.floor(v)) - 1.0)
            del i, v
            dst, xo, rhs, dinv = (ctx.alloc(a.nb * N) for _ in range(4))
            ctx.call("spirk_op_inverse_diagonal", C.byref(lvl), dinv, 16.0, 0.1)
            op = capi.real_op([16.0, 3.16, 2.94, 5.64, 1.0, 2.0, 3.0, 4.0][:a.nb], [0.1])
            f1, _1 = capi.darr([0.3] * a.nb)
            f2, _2 = capi.darr([1.1] * a.nb)
            for o in a.opt:
                name, value = o.split("=")
                ctx.call("spirk_ctx_set_option", name.encode(), int(value))
            for variant in a.variants:
                ctx.call("spirk_ctx_set_option", b"apply_variant", variant)
                for name, fn, bytes_per_dof in [
                    ("apply", lambda: ctx.call("spirk_op_apply", C.byref(lvl), C.byref(op), dst, src, N), 16),
                    ("cheb_step", lambda: ctx.call("spirk_op_cheb_step", C.byref(lvl), C.byref(op), dst, src, xo, rhs,
                                                   dinv, N, f1, f2), 40),
                    ("cheb_step_own_dinv", lambda: ctx.call("spirk_op_cheb_step", C.byref(lvl), C.byref(op), dst, src, xo,
                                                            rhs, None, N, f1, f2), 32),
                    ("residual", lambda: ctx.call("spirk_op_residual", C.byref(lvl), C.byref(op), dst, rhs, src, N), 24),
                ]:
                    if a.kernels and name not in a.kernels:
                        continue
                    for _ in range(3):
                        fn()
                    ctx.call("spirk_ctx_timer_begin")
                    for _ in range(a.reps):
                        fn()
                    ms = ctx.scalar_call("spirk_ctx_timer_end") / a.reps
                    gdofs = a.nb * N / ms * 1e-6
                    gbs = gdofs * bytes_per_dof
                    print(json.dumps({"tag": a.tag, "opt": a.opt, "kernel": name, "variant": variant, "dim": a.dim, "degree": a.degree, "refine": r,
                                      "nb": a.nb, "n_dofs": N, "ms": round(ms, 4), "gdof_per_s": round(gdofs, 2),
                                      "algorithmic_gb_per_s": round(gbs, 1),
                                      "frac_of_measured_hbm": round(gbs / peaks["hbm_gbs"], 3)}), flush=True)


if __name__ == "__main__":
    main()
