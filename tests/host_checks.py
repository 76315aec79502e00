"""End-to-end checks of the C++ host layer (reference API mirror) against the NumPy oracle.

Run in the CPU suite with the host layer linked against the CPU double, and in the `-m gpu` suite
with the product libraries.  Bars (north_star): iteration counts identical +-1, solution and error
norms within 1e-10 relative (parity runs use OuterTolerance 1e-12)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import spirk_oracle as so  # noqa: E402
from dealii_spirk_b200 import hostapi  # noqa: E402

TABLES = os.path.join(ROOT, "dealii_spirk_b200", "tables", "butcher_tables.txt")


def params(scheme, k, r, q, tol=1e-12, inner=0.0, tau=0.1, end=0.5, **extra):
    p = {"FEDegree": k, "NRefinements": r, "TimeIntegrationScheme": scheme, "IRKStages": q, "TimeStepSize": tau,
         "EndTime": end, "OperatorType": "MatrixFree", "BlockPreconditionerType": "GMG", "OuterTolerance": tol,
         "InnerTolerance": inner, "DoOutputParaview": False}
    p.update(extra)
    return p


def run_host(host, scheme, dim, k, r, q, **kw):
    with hostapi.Run(host, params(scheme, k, r, q, **kw), dim=dim) as run:
        run.setup()
        while not run.finished():
            run.step()
        run.finish()
        return {"u": run.solution(), "outer": run.array("outer_iterations"), "inner": run.array("inner_iterations"),
                "error_L2": run.array("error_L2"), "error_Linf": run.array("error_Linf"), "norm": run.array("solution_l2"),
                "table": run.table_text()}


def compare(host, scheme, dim, k, r, q, tol=1e-12, inner=0.0, sol_tol=1e-10, count_slack=1, **okw):
    res = run_host(host, scheme, dim, k, r, q, tol=tol, inner=inner)
    ora = so.run(scheme, dim, k, r, q, 0.1, 0.5, outer_tol=tol, inner_tol=inner, **okw)
    it = ora["integ"]
    uo = ora["u"].reshape(-1)
    rel = np.max(np.abs(res["u"] - uo)) / np.max(np.abs(uo))
    assert rel < sol_tol, f"{scheme}: solution differs from the oracle by {rel}"
    eo = np.array(ora["errors"])
    assert np.allclose(res["error_L2"], eo[:, 0], rtol=1e-7, atol=0), (res["error_L2"], eo[:, 0])
    assert np.allclose(res["error_Linf"], eo[:, 1], rtol=1e-7, atol=0)
    assert np.allclose(res["norm"][1:], ora["norms"], rtol=1e-10, atol=0)
    if scheme == "ost":
        pass
    elif scheme.startswith("complex"):
        ref = np.array([max(o) for o in it.n_outer])
        assert np.all(np.abs(res["outer"] - ref) <= count_slack), (res["outer"], it.n_outer)
    else:
        assert np.all(np.abs(res["outer"] - np.array(it.n_outer)) <= count_slack), (res["outer"], it.n_outer)
        assert np.all(np.abs(res["inner"] - np.array(it.n_inner)) <= count_slack * q), (res["inner"], it.n_inner)
    return res, ora


def oracle_gmg_iterations(dim, k, r, mode, n_components):
    """CG iteration count of the reference's gmg.cc benchmark (gmg.cc:212-278) from the NumPy oracle:
    (M + K) u = 1, ReductionControl(1000, 1e-20, 1e-12), one GMG V-cycle as preconditioner."""
    import numpy as np
    import spirk_oracle as so
    lv = so.Level(dim, k, r)
    control = so.SolverControl(1000, 1e-20, reduce=1e-12)
    counts = []
    orig_check = control.check

    def check(step, value):
        counts.append(step)
        return orig_check(step, value)

    control.check = check
    if mode in (0, 2):
        op = so.ScalarOp(lv, 1.0, 1.0)
        gmg = so.GMG(dim, k, r, lambda l: so.ScalarOp(l, 1.0, 1.0))
        gmg.reinit()
        so.solver_cg(op.vmult, lv.zeros(), np.ones((1,) + lv.shape), gmg.vmult, control)
    else:
        d_vec = np.ones(n_components)
        op = so.BatchedOp(lv, d_vec, 1.0)
        b = np.ones((n_components,) + lv.shape)
        if mode == 1:  # the scalar V-cycle on every component
            gmg = so.GMG(dim, k, r, lambda l: so.ScalarOp(l, 1.0, 1.0))
            gmg.reinit()
            P = lambda g: np.concatenate([gmg.vmult(g[i:i + 1]) for i in range(n_components)])  # noqa: E731
        else:          # block GMG (one Chebyshev range, smoother as coarse solver)
            gmg = so.GMG(dim, k, r, lambda l: so.BatchedOp(l, d_vec, 1.0), block=True)
            gmg.reinit()
            P = gmg.vmult
        so.solver_cg(op.vmult, np.zeros_like(b), b, P, control)
    return counts[-1]


def check_gmg_benchmark(host, dim, k, r, n_components=3, slack=1):
    """all four modes of the gmg.cc benchmark: table columns and iteration counts against the oracle"""
    rows = {}
    for mode in range(4):
        row = host.gmg(dim, k, r, mode, n_components=n_components, n_repetitions=1)
        ref = oracle_gmg_iterations(dim, k, r, mode, n_components)
        assert abs(row["n_iterations"] - ref) <= slack, (mode, row, ref)
        assert row["L"] == r + 1 and row["n_cells"] == 2 ** (dim * r) and row["time"] > 0.0
        n1 = k * 2 ** r + 1
        assert row["n_dofs"] == n1 ** dim * (n_components if mode == 1 else 1)
        rows[mode] = row
    assert rows[0]["n_iterations"] == rows[2]["n_iterations"]
    return rows
