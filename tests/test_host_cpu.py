"""CPU suite: the C++ host layer (reference API mirror: operators, GMG, CG/GMRES, all eight time
integration schemes) linked against the CPU double of the C ABI, checked against the independent
NumPy oracle: solution 1e-10, iteration counts +-1 (north_star's parity bar)."""
import os
import subprocess

import pytest

import host_checks as hc
from dealii_spirk_b200 import capi, hostapi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cpu_host():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"])
    return hostapi.HostLib(os.path.join(ROOT, "oracle", "_build", "libspirk_host_cpu.so"), hc.TABLES)


CASES = [
    ("irk", 2, 2, 3, 2), ("irk", 3, 4, 2, 4), ("spirk", 3, 4, 1, 2), ("spirk", 2, 2, 3, 4), ("irk_batched", 3, 4, 2, 4),
    ("complex_irk", 3, 4, 2, 4), ("complex_irk_batched", 3, 4, 1, 4), ("complex_spirk", 2, 2, 3, 3),
    ("complex_spirk_batched", 2, 2, 3, 2), ("ost", 2, 2, 4, 0), ("irk", 3, 1, 3, 2), ("irk", 2, 4, 2, 5),
]


@pytest.mark.parametrize("scheme,dim,k,r,q", CASES)
def test_scheme_matches_oracle(cpu_host, scheme, dim, k, r, q):
    assert cpu_host.backend() == "cpu-oracle"
    hc.compare(cpu_host, scheme, dim, k, r, q)


def test_default_outer_tolerance_counts(cpu_host):
    # 3-D Q4 r=2 q=4 at the reference's default OuterTolerance 1e-8 -> 7 7 7 6 7
    res, _ = hc.compare(cpu_host, "irk", 3, 4, 2, 4, tol=1e-8, sol_tol=1e-6)
    assert list(res["outer"]) == [7, 7, 7, 6, 7]


def test_inner_tolerance_path(cpu_host):
    # InnerTolerance > 0: inner CG per stage (main.cc:1126-1141); SURVEY 8f rank 1
    hc.compare(cpu_host, "irk", 2, 2, 3, 2, inner=1e-6, count_slack=1)
    hc.compare(cpu_host, "complex_irk", 2, 2, 3, 2, inner=1e-6, count_slack=1)


def test_parameter_errors(cpu_host):
    with pytest.raises(capi.SpirkError):
        hostapi.Run(cpu_host, hc.params("bogus", 2, 2, 2))
    with pytest.raises(capi.SpirkError):
        run = hostapi.Run(cpu_host, hc.params("irk", 2, 2, 2, OperatorType="MatrixBased"), dim=2)
        run.setup()
    with pytest.raises(capi.SpirkError):  # no complex tables for q = 10
        run = hostapi.Run(cpu_host, hc.params("complex_irk", 1, 2, 10), dim=2)
        run.setup()


def test_gmg_benchmark_modes(cpu_host):
    # the reference's gmg.cc benchmark (1 component / n components / n groups / batched)
    hc.check_gmg_benchmark(cpu_host, 2, 2, 3)
    hc.check_gmg_benchmark(cpu_host, 3, 4, 1, n_components=2)


def test_dealii_renumbering_glue(tmp_path):
    """INTEGRATION.md route A: dealii_spirk_b200/host/dealii_glue.h maps a (mock) deal.II support-point map to the library's
    lexicographic numbering for every degree, 2-D and 3-D (reference DoF distribution: main.cc:3374-3412)"""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "glue_check")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(root, "tests", "glue_check.cc")])
    assert "glue ok" in subprocess.run([exe], capture_output=True, text=True, check=True).stdout
