#!/usr/bin/env python3
"""The reference's gmg.cc benchmark (GMG-preconditioned CG on (M + K) u = 1, time per CG iteration) in its four
modes: one component, n components in one system, one component per process group, n components batched
(gmg.cc:342-382).  Prints the reference's table columns; `--cpu` runs the CPU double (test infrastructure)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dealii_spirk_b200 as pkg  # noqa: E402
from dealii_spirk_b200 import hostapi  # noqa: E402

MODES = {0: "1 component", 1: "n components (one system)", 2: "n groups x 1 component", 3: "n components batched"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=3)
    ap.add_argument("--degree", type=int, default=4)  # the reference's default is 1; Q4 is the path's degree
    ap.add_argument("--min-refine", type=int, default=3)
    ap.add_argument("--max-refine", type=int, default=6)
    ap.add_argument("--components", type=int, default=8)
    ap.add_argument("--repetitions", type=int, default=10)
    ap.add_argument("--modes", type=int, nargs="+", default=[0, 1, 2, 3])
    ap.add_argument("--cpu", action="store_true")
    a = ap.parse_args()
    root = os.path.dirname(pkg.HERE)
    if a.cpu:
        host = hostapi.HostLib(os.path.join(root, "oracle", "_build", "libspirk_host_cpu.so"), pkg.TABLES_PATH)
    else:
        pkg.device_lib()
        host = hostapi.HostLib(pkg.HOST_LIB_PATH, pkg.TABLES_PATH)
    print(f"backend {host.backend()}")
    print(" ".join(f"{c:>13s}" for c in hostapi.HostLib.GMG_COLUMNS + ("GDoF/s/it", "mode")))
    for r in range(a.min_refine, a.max_refine + 1):
        for mode in a.modes:
            row = host.gmg(a.dim, a.degree, r, mode, a.components, a.repetitions)
            n_unknowns = row["n_dofs"] * (a.components if mode == 3 else 1)
            print(" ".join(f"{row[c]:13.6g}" for c in hostapi.HostLib.GMG_COLUMNS),
                  f"{n_unknowns / row['time'] * 1e-9:13.3f}", MODES[mode], flush=True)
            print(json.dumps({"gmg": row, "mode": mode, "refine": r}), file=sys.stderr)


if __name__ == "__main__":
    main()
