"""Golden fixtures (tests/golden/): CPU suite = oracle + CPU double; -m gpu = CUDA path through the C ABI."""
import json
import os
import subprocess

import numpy as np
import pytest

import golden_checks as gc
import host_checks as hc
import spirk_oracle as so
from dealii_spirk_b200 import hostapi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_oracle_reproduces_direct_golden():
    direct = json.load(open(os.path.join(gc.GOLD, "direct_solve.json")))
    rec = direct["2d_q2_r3_s2"]
    prob = so.Problem(2, 2, 3)
    u, t = prob.initial(), 0.0
    for s in range(5):
        t += 0.1
        u, _ = so.direct_irk_step(prob, 2, 0.1, u, t)
        assert abs(np.sqrt(so.dot(u, u)) - rec["steps"][s]["l2_norm"]) < 1e-12 * rec["steps"][s]["l2_norm"]
    assert np.allclose(u.reshape(-1), rec["u_final"], rtol=0, atol=1e-13)
    # SURVEY Appendix D row "2-D Q2 r=3 q=2"
    assert abs(rec["steps"][0]["error_L2"] - 2.419429e-03) < 1e-8 and abs(rec["steps"][4]["l2_norm"] - 12.46545730316) < 1e-10


def test_cpu_double_matches_golden_operator(cpu_dev):
    gc.check_operator_outputs(cpu_dev)


def test_cpu_host_matches_golden_runs():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"])
    host = hostapi.HostLib(os.path.join(ROOT, "oracle", "_build", "libspirk_host_cpu.so"), hc.TABLES)
    gc.check_iterative_runs(host, keys=["irk_2d_q2_r3_s2", "complex_irk_2d_q2_r3_s3", "ost_2d_q2_r4_s0"])


@pytest.mark.gpu
def test_gpu_matches_golden_operator(gpu_dev):
    gc.check_operator_outputs(gpu_dev)


@pytest.mark.gpu
def test_gpu_host_matches_golden_runs(gpu_dev):
    import dealii_spirk_b200 as pkg
    host = hostapi.HostLib(pkg.HOST_LIB_PATH, hc.TABLES)
    gc.check_iterative_runs(host)
