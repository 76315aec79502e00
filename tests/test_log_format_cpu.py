"""CPU suite: the printed per-step lines and the statistics table have the reference's format, and the
automatic time step follows the reference's rule (main.cc:3314-3318)."""
import json
import os
import re
import subprocess
import sys

import pytest

import host_checks as hc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLOAT = r"[-+]?\d+(\.\d+)?(e[-+]?\d+)?"

# the columns of the reference's ConvergenceTable in the order they are added: Problem::run main.cc:3387-3398, IRKBase::get_statistics
# 695-718 + 810-812, Problem::run 3360-3368 (the table prints the columns in order of first appearance)
TABLE_COLUMNS = ["n_levels", "n_cells", "fe_degree", "n_dofs", "n_stages", "n_procs", "n_procs_global", "n_procs_row",
                 "n_procs_column", "n_t", "final_t", "dt", "error_L2", "error_Linf", "n_outer_min", "n_outer_avg", "n_outer_max",
                 "n_inner_min", "n_inner_avg", "n_inner_max", "t", "t_rhs", "t_solver", "t_update", "t_vmult", "t_prec_bc",
                 "t_prec_solver"] + [f"t_prec_solver_{i}" for i in range(10)]


def run_verbose(scheme, dim, k, r, q, **kw):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"])
    lib = os.path.join(ROOT, "oracle", "_build", "libspirk_host_cpu.so")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "log_worker.py"), lib, hc.TABLES, str(dim),
                        json.dumps(hc.params(scheme, k, r, q, **kw))], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, OMP_NUM_THREADS="4"))
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout


def check_common(out, n_steps):
    assert re.search(r"^Number of active cells: \d+$", out, re.M)
    assert re.search(r"^Number of degrees of freedom: \d+$", out, re.M)
    assert re.search(rf"^Starting time loop with dt={FLOAT}$", out, re.M)
    steps = re.findall(rf"^Time step (\d+) at t=({FLOAT})$", out, re.M)
    assert [int(s[0]) for s in steps] == list(range(n_steps)), steps     # the reference prints the number BEFORE incrementing it
    errs = re.findall(rf"^   Error in the L2/L∞ norm : {FLOAT}/{FLOAT}$", out, re.M)
    assert len(errs) == n_steps + 1                                      # initial value + every step


def check_table(out, columns):
    table = out.split("TABLE_TEXT_BEGIN")[1].strip().splitlines()
    header = table[0].split()
    assert header == columns, (header, columns)
    row = table[1].split()
    assert len(row) == len(columns)
    for name, v in zip(columns, row):
        float(v)
        if name.startswith("t") or name in ("final_t", "dt", "error_L2", "error_Linf"):
            assert re.fullmatch(r"[-+]?\d\.\d+e[-+]\d+", v), (name, v)   # set_scientific columns


def test_irk_lines_and_table():
    out = run_verbose("irk", 2, 2, 3, 2, tol=1e-8, end=0.3)
    check_common(out, 3)
    lines = re.findall(r"^   (\d+) outer GMRES iterations and (\d+)\+(\d+) inner CG iterations\.$", out, re.M)
    assert len(lines) == 3 and all(int(a) > 0 and int(b) == int(a) + 1 for a, b, _ in lines)  # one V-cycle per (re)start + iteration
    check_table(out, TABLE_COLUMNS)


def test_irk_batched_and_spirk_lines():
    out = run_verbose("irk_batched", 2, 2, 3, 2, tol=1e-8, end=0.2)
    assert len(re.findall(r"^   \d+ outer GMRES iterations and \d+ inner CG iterations\.$", out, re.M)) == 2
    out = run_verbose("spirk", 2, 2, 3, 2, tol=1e-8, end=0.2)  # one rank: the IRK line
    assert len(re.findall(r"^   \d+ outer GMRES iterations and \d+\+\d+ inner CG iterations\.$", out, re.M)) == 2


def test_complex_lines_and_table():
    out = run_verbose("complex_irk", 2, 2, 3, 4, tol=1e-8, end=0.2)
    check_common(out, 2)
    assert len(re.findall(r"^   Solved in: \d+ \(\d+\+\d+\), \d+ \(\d+\+\d+\)$", out, re.M)) == 2
    check_table(out, TABLE_COLUMNS)
    out = run_verbose("complex_irk_batched", 2, 2, 3, 3, tol=1e-8, end=0.2)
    assert len(re.findall(r"^   Solved in: \d+ \(\d+\), \d+ \(\d+\)$", out, re.M)) == 2


def test_ost_lines_and_table():
    out = run_verbose("ost", 2, 2, 3, 0, end=0.2)
    check_common(out, 2)
    assert len(re.findall(r"^   \d+ CG iterations\.$", out, re.M)) == 2
    cols = [c for c in TABLE_COLUMNS if not c.startswith(("n_inner", "t")) and c not in ("n_outer_min", "n_outer_max")] + []
    table = out.split("TABLE_TEXT_BEGIN")[1].strip().splitlines()
    assert table[0].split() == ["n_levels", "n_cells", "fe_degree", "n_dofs", "n_stages", "n_procs", "n_procs_global", "n_procs_row",
                                "n_procs_column", "n_t", "final_t", "dt", "error_L2", "error_Linf", "n_outer_avg"], (table[0], cols)


@pytest.mark.parametrize("k,r,q", [(2, 3, 2), (1, 4, 3)])
def test_automatic_time_step(k, r, q):
    """TimeStepSize <= 0: dt = dx^((k+1)/(2q-1)), dx = minimum vertex distance = 2^-r (main.cc:3309-3318)"""
    out = run_verbose("irk", 2, k, r, q, tol=1e-8, tau=0.0, end=0.8)
    dt = float(re.search(r"^DT=(.*)$", out, re.M).group(1))
    want = (2.0 ** -r) ** ((k + 1.0) / (2.0 * q - 1.0))
    assert abs(dt - want) < 1e-14 * want
    assert re.search(rf"^Starting time loop with dt={FLOAT}$", out, re.M)
