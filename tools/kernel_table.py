#!/usr/bin/env python3
"""Per-kernel table from an ncu CSV with several metrics per launch (gpu__time_duration.sum, dram__bytes_read.sum,
dram__bytes_write.sum, FP64 pipe / L2 hit rate / warps active / issue active): launches grouped by (kernel, grid size), i.e.
by multigrid level; time, measured DRAM bytes and achieved DRAM GB/s per launch.  With --dofs N the rows whose algorithmic
bytes per DoF are known (ALGO below) also show algorithmic GB/s and its fraction of the measured copy peak.
Usage: kernel_table.py X.csv [--dofs 33949186] [--peak 6557.4] [--min-us 15]"""
import argparse
import csv
import re
from collections import defaultdict

# algorithmic bytes per stage-DoF of the kernels the solver spends its time in
ALGO = {"k_v3<4, 8, 8, 0, 2, 1>": 16, "k_v3<4, 8, 8, 1, 2, 1>": 24, "k_v3<4, 8, 8, 3, 2, 1>": 32, "k_v3<4, 8, 8, 4, 2, 1>": 24,
        "k_v3<4, 8, 8, 0, 4, 2>": 16, "k_v3<4, 8, 8, 1, 4, 2>": 24, "k_v3<4, 8, 8, 3, 4, 2>": 32, "k_v3<4, 8, 8, 4, 4, 2>": 24,
        "k_mix<2>": 16, "k_sub_and_dot_dev": 24, "k_axpy": 24, "k_dot": 16, "k_scale": 16, "k_add_and_dot": 24}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--dofs", type=float, default=0.0, help="stage-DoFs per launch of the finest level (n_dofs * blocks)")
    ap.add_argument("--peak", type=float, default=6557.4)
    ap.add_argument("--min-us", type=float, default=15.0)
    a = ap.parse_args()
    rows = [r for r in csv.reader(l for l in open(a.csv) if l.startswith('"'))]
    head, rows = rows[0], rows[1:]
    iid, ik, ig, im, iu, iv = (head.index(k) for k in ("ID", "Kernel Name", "Grid Size", "Metric Name", "Metric Unit", "Metric Value"))
    launches = defaultdict(dict)
    names = {}
    for r in rows:
        v = float(r[iv].replace(",", ""))
        u = r[iu]
        if r[im].startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        if r[im] == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(u, 1.0)
        launches[r[iid]][r[im]] = v
        names[r[iid]] = (re.sub(r"\(.*$", "", re.sub(r"^(void )?(spirk::)?", "", r[ik])), r[ig])
    groups = defaultdict(list)
    for i, m in launches.items():
        groups[names[i]].append(m)
    print(f"{'kernel':34s} {'grid':>14s} {'n':>4s} {'us':>8s} {'MB rd':>8s} {'MB wr':>8s} {'DRAM GB/s':>9s} {'fp64%':>6s} {'L2hit%':>6s} {'warps%':>6s} {'issue%':>6s}  algorithmic")
    tot = sum(m.get("gpu__time_duration.sum", 0.0) for g in groups.values() for m in g)
    for (k, g), ms in sorted(groups.items(), key=lambda kv: -sum(m.get("gpu__time_duration.sum", 0.0) for m in kv[1])):
        avg = lambda key: sum(m.get(key, 0.0) for m in ms) / len(ms)
        t = avg("gpu__time_duration.sum")
        if t < a.min_us:
            continue
        rd, wr = avg("dram__bytes_read.sum"), avg("dram__bytes_write.sum")
        extra = ""
        if a.dofs and k in ALGO and t > 0.5 * max(avg_t for (kk, _), mm in groups.items() if kk == k for avg_t in [sum(x.get("gpu__time_duration.sum", 0) for x in mm) / len(mm)]):
            gbs = ALGO[k] * a.dofs / t * 1e-3
            extra = f"{ALGO[k]} B/DoF: {gbs:7.0f} GB/s = {gbs / a.peak:.2f} of peak; DRAM/algorithmic {(rd + wr) / (ALGO[k] * a.dofs):.2f}"
        print(f"{k:34s} {g:>14s} {len(ms):4d} {t:8.1f} {rd / 1e6:8.1f} {wr / 1e6:8.1f} {(rd + wr) / t * 1e-3:9.0f} "
              f"{avg('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'):6.1f} {avg('lts__t_sector_hit_rate.pct'):6.1f} "
              f"{avg('sm__warps_active.avg.pct_of_peak_sustained_active'):6.1f} {avg('smsp__issue_active.avg.pct_of_peak_sustained_active'):6.1f}  {extra}")
    print(f"total {tot / 1e3:.2f} ms in {len(launches)} launches (serialised, per-launch cold caches)")


if __name__ == "__main__":
    main()
