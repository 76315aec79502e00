#!/usr/bin/env python3
"""Generate tests/golden/*.json from the independent NumPy/SciPy oracle (oracle/spirk_oracle.py).

The reference ships no golden vectors and cannot be built or imported here, so the
fixtures pin (a) the known-answer values of SURVEY Appendix D re-derived by a sparse-direct solve,
(b) cell-operator outputs on the stateless synthetic input of SURVEY 8(d), (c) iteration counts and
final solutions of the iterative path.  Run in the build container:  python tools/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import spirk_oracle as so  # noqa: E402
from abi_checks import OP_CASES, block_input  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    # (a) direct solves
    direct = {}
    for (dim, k, r, q) in [(2, 2, 3, 2), (2, 4, 3, 4), (3, 1, 3, 2), (3, 4, 2, 4), (3, 4, 2, 2)]:
        prob = so.Problem(dim, k, r)
        u, t, steps = prob.initial(), 0.0, []
        e0 = prob.errors(u, 0.0)
        for s in range(5):
            t += 0.1
            u, _ = so.direct_irk_step(prob, q, 0.1, u, t)
            e = prob.errors(u, t)
            steps.append({"t": t, "error_L2": e[0], "error_Linf": e[1], "l2_norm": float(np.sqrt(so.dot(u, u)))})
        direct[f"{dim}d_q{k}_r{r}_s{q}"] = {"dim": dim, "k": k, "r": r, "q": q, "error_t0": list(e0), "steps": steps,
                                             "u_final": u.reshape(-1).tolist() if u.size <= 1200 else None}
    json.dump(direct, open(os.path.join(OUT, "direct_solve.json"), "w"), indent=0)
    # (b) operator outputs, 3-D Q4 r=1 (729 DoFs) and 2-D Q2 r=2 (81 DoFs)
    ops = []
    for (dim, k, r) in [(3, 4, 1), (2, 2, 2)]:
        lv = so.Level(dim, k, r)
        for case in OP_CASES:
            if case[0] == "real":
                mass = np.atleast_1d(np.asarray(case[1], float))
                lap = np.broadcast_to(np.atleast_1d(np.asarray(case[2], float)), mass.shape)
                u = block_input(lv, len(mass), seed=1)
                ref = lv.apply(u, mass, lap)
            else:
                Cm = np.asarray(case[1], float)
                nb = Cm.shape[0]
                lap = np.broadcast_to(np.atleast_1d(np.asarray(case[2], float)), (nb,))
                u = block_input(lv, nb, seed=2)
                v = u.copy()
                v[:, lv.bmask] = 0.0
                ref = lv.apply(v, 0.0, lap) + np.tensordot(Cm, lv.apply(v, 1.0, 0.0), axes=(1, 0))
                ref[:, lv.bmask] = u[:, lv.bmask]
            ops.append({"dim": dim, "k": k, "r": r, "case": [case[0], np.asarray(case[1]).tolist(), list(case[2])],
                        "out": ref.reshape(-1).tolist()})
    json.dump(ops, open(os.path.join(OUT, "operator_outputs.json"), "w"))
    # (c) iterative path: iteration counts + solution of every scheme on small meshes (OuterTolerance 1e-12)
    it = {}
    for (scheme, dim, k, r, q) in [("irk", 2, 2, 3, 2), ("irk", 3, 4, 2, 4), ("irk_batched", 3, 4, 2, 4), ("complex_irk", 3, 4, 2, 4),
                                   ("complex_irk_batched", 3, 4, 2, 4), ("complex_irk", 2, 2, 3, 3), ("ost", 2, 2, 4, 0)]:
        o = so.run(scheme, dim, k, r, q, 0.1, 0.5, outer_tol=1e-12)
        integ = o["integ"]
        counts = integ.n_iter if scheme == "ost" else integ.n_outer
        it[f"{scheme}_{dim}d_q{k}_r{r}_s{q}"] = {"scheme": scheme, "dim": dim, "k": k, "r": r, "q": q, "outer": counts,
                                                 "errors": o["errors"], "norms": o["norms"],
                                                 "u_final": o["u"].reshape(-1).tolist() if o["u"].size <= 5000 else None}
    json.dump(it, open(os.path.join(OUT, "iterative_runs.json"), "w"))
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
