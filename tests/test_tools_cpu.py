"""The profile post-processing tools reproduce the committed summaries from the committed ncu CSV files (no GPU needed):
the kernel shares / per-kernel DRAM figures quoted in DESIGN.md can be re-derived by anyone."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")


def run_tool(*args):
    return subprocess.run([sys.executable] + list(args), capture_output=True, text=True, check=True, cwd=ROOT).stdout


def test_launch_list_summary_matches_committed_summary():
    out = run_tool("tools/summarize_launches.py", os.path.join(PROF, "r02_launches_bench_r6_final.csv"), "title")
    committed = open(os.path.join(PROF, "r02_launches_bench_r6_final_summary.txt")).read()
    # same table below the title line
    assert out.split("\n", 1)[1] == committed.split("\n", 1)[1]
    m = re.search(r"total ([0-9.]+) ms over (\d+) launches", out)
    assert m and int(m.group(2)) == 1600
    top = out.splitlines()[3]
    assert "k_v3<4, 8, 8, 3, 2, 1>" in top  # the fused Chebyshev step dominates the step


def test_kernel_table_from_metrics_csv():
    out = run_tool("tools/kernel_table.py", os.path.join(PROF, "r02_kernels_bench_r6_metrics.csv"), "--dofs", "33949186")
    rows = {(l.split("(")[0].strip(), l[l.index("("):l.index(")") + 1]): l for l in out.splitlines() if l.startswith("k_")}
    cheb = rows[("k_v3<4, 8, 8, 3, 2, 1>", "(296, 1, 1)")]
    # algorithmic 32 B per stage-DoF; measured DRAM traffic within 10 % of it (no wasted re-reads)
    ratio = float(re.search(r"DRAM/algorithmic ([0-9.]+)", cheb).group(1))
    assert 0.9 < ratio < 1.1, cheb
    frac = float(re.search(r"= ([0-9.]+) of peak", cheb).group(1))
    assert 0.4 < frac < 0.6, cheb
    # the coupled pair stages the plane of both blocks for every output block: more traffic than algorithmic
    pair = rows[("k_v3<4, 8, 8, 0, 4, 2>", "(296, 1, 1)")]
    assert float(re.search(r"DRAM/algorithmic ([0-9.]+)", pair).group(1)) > 1.3
    assert "total" in out.splitlines()[-1]
