"""-m gpu: the product stack (CUDA kernels + C++ host layer) through the host C API, against the
NumPy oracle.  Same bars as the CPU suite: solution / error norms 1e-10, iteration counts +-1."""
import numpy as np
import pytest

import host_checks as hc
from dealii_spirk_b200 import hostapi
import dealii_spirk_b200 as pkg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu_host():
    pkg.device_lib()
    host = hostapi.HostLib(pkg.HOST_LIB_PATH, hc.TABLES)
    assert host.backend() == "cuda-sm_100a"
    return host


CASES = [
    ("irk", 2, 2, 3, 2), ("irk", 3, 4, 2, 4), ("irk", 3, 4, 3, 2), ("spirk", 3, 4, 2, 4), ("irk_batched", 3, 4, 2, 4),
    ("complex_irk", 3, 4, 2, 4), ("complex_irk_batched", 3, 4, 2, 4), ("complex_spirk", 2, 2, 3, 3),
    ("complex_spirk_batched", 3, 4, 1, 8), ("ost", 2, 2, 5, 0), ("irk", 3, 1, 3, 2), ("spirk", 3, 4, 2, 8),
]


@pytest.mark.parametrize("scheme,dim,k,r,q", CASES)
def test_scheme_matches_oracle(gpu_host, scheme, dim, k, r, q):
    # q = 8: attainable accuracy of the stage vectors is ~1e-10 (cond(T) = 7e5)
    hc.compare(gpu_host, scheme, dim, k, r, q, sol_tol=1e-10 if q < 8 else 1e-9)


def test_inner_tolerance_path(gpu_host):
    hc.compare(gpu_host, "irk", 2, 2, 3, 2, inner=1e-6)


def test_end_to_end_host_buffers(gpu_host):
    """the reference-facing call with HOST buffers: Interface::solve on a host solution vector"""
    p = hc.params("irk", 4, 2, 2)
    with hostapi.Run(gpu_host, p, dim=3) as a, hostapi.Run(gpu_host, p, dim=3) as b:
        a.setup(), b.setup()
        u = b.solution()
        for _ in range(3):
            a.step()
            b.step_host(u)
        ua = a.solution()
        assert np.max(np.abs(ua - u)) <= 1e-11 * np.max(np.abs(u))


def test_manufactured_solution_full_size(gpu_host):
    """size-independent property up to a BASELINE size (3-D Q4 r=5, 2.1e6 DoFs x 4 stages): with q=4
    (time error ~ tau^7) the L2 error against the analytical solution is the spatial error and falls
    like h^(k+1) = 1/32 per refinement; the outer iteration count stays mesh independent."""
    errs = {}
    for r in (3, 4, 5):
        res = hc.run_host(gpu_host, "irk", 3, 4, r, 4, tol=1e-10, end=0.2)
        errs[r] = res["error_L2"][-1]
        assert np.all(res["outer"] <= 10)
    # r=3 -> r=4: spatial order 5 (factor 32); at r=5 the tau^7 time error (~1e-8) starts to show
    assert errs[4] < errs[3] / 16.0 and errs[5] < errs[4] / 4.0, errs


def test_gmg_benchmark_modes(gpu_host):
    """the reference's gmg.cc benchmark modes on the GPU: iteration counts as the oracle"""
    hc.check_gmg_benchmark(gpu_host, 3, 4, 2, n_components=3)
    hc.check_gmg_benchmark(gpu_host, 2, 2, 4, n_components=8)


def test_baseline_size_against_oracle_fixture(gpu_host):
    """BASELINE configs[1] at r = 5 (3-D Q4, IRK q = 2, 2 146 689 DoFs x 2 stages): the CUDA path against the NumPy oracle's
    answer committed in tests/golden/baseline_size_irk_q2_r5.json (tools/make_golden_baseline_size.py; the oracle needs
    minutes for this size): solution 1e-10, error norms, solution norm, outer iteration counts +-1."""
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "baseline_size_irk_q2_r5.json")))
    res = hc.run_host(gpu_host, g["scheme"], g["dim"], g["k"], g["r"], g["q"], tol=g["outer_tol"], end=g["end"])
    us = res["u"][::g["sample_stride"]]
    assert np.max(np.abs(us - np.array(g["u_sample"]))) < 1e-10 * g["u_max"]
    eo = np.array(g["errors"])
    assert np.allclose(res["error_L2"], eo[:, 0], rtol=1e-7, atol=0) and np.allclose(res["error_Linf"], eo[:, 1], rtol=1e-7, atol=0)
    assert np.allclose(res["norm"][1:], g["norms"], rtol=1e-10, atol=0)
    assert np.all(np.abs(res["outer"] - np.array(g["n_outer"])) <= 1), (res["outer"], g["n_outer"])
