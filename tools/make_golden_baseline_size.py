#!/usr/bin/env python3
"""tests/golden/baseline_size_irk_q2_r5.json: the NumPy oracle's answer for BASELINE configs[1] at r = 5 (3-D Q4, IRK q = 2,
2 146 689 DoFs x 2 stages, OuterTolerance 1e-12, two time steps): error norms, solution norm, outer iteration counts and the
solution at every 4099-th DoF.  The oracle needs minutes for this on a CPU, so the answer is committed as a fixture and the
-m gpu suite compares the CUDA path with it (tests/test_gpu_host.py).  Run in the build container:
    python tools/make_golden_baseline_size.py
The reference itself pins nothing and cannot be built here: this is an oracle fixture, not reference output."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import spirk_oracle as so  # noqa: E402

if __name__ == "__main__":
    t0 = time.time()
    dim, k, r, q, end = 3, 4, 5, 2, 0.2
    ora = so.run("irk", dim, k, r, q, 0.1, end, outer_tol=1e-12)
    u = ora["u"].reshape(-1)
    out = {"scheme": "irk", "dim": dim, "k": k, "r": r, "q": q, "tau": 0.1, "end": end, "outer_tol": 1e-12,
           "n_outer": [int(x) for x in ora["integ"].n_outer], "errors": [[float(a), float(b)] for a, b in ora["errors"]],
           "norms": [float(x) for x in ora["norms"]], "sample_stride": 4099, "u_sample": u[::4099].tolist(),
           "u_max": float(np.max(np.abs(u))), "seconds": time.time() - t0}
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "baseline_size_irk_q2_r5.json"), "w"))
    print("done in", time.time() - t0, "s:", out["n_outer"], out["errors"][-1])
