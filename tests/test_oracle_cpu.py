"""CPU suite: the two oracles against each other and against the known-answer values of
SURVEY Appendix D; the product library's symbol table (no compute without a GPU)."""
import ctypes as C
import os

import numpy as np
import pytest

import abi_checks as ac
import spirk_oracle as so
from dealii_spirk_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_table_identities():
    # SURVEY 8c(2)
    for q in range(2, 10):
        A, Ai, T, Ti = so.table("A", q), so.table("A_inv", q), so.table("T", q), so.table("T_inv", q)
        D, L, b, c = so.table("D_vec_", q), so.table("L", q), so.table("b_vec_", q), so.table("c_vec_", q)
        assert np.abs(A @ Ai - np.eye(q)).max() < 1e-12
        assert np.abs(T @ np.diag(D) @ Ti - L).max() < 1e-9 * max(1.0, np.abs(L).max())
        assert np.abs(b - A[-1]).max() < 1e-6 and abs(c[-1] - 1.0) < 1e-14
        assert np.count_nonzero(np.abs(T) > 1e-12) == q * (q + 1) // 2
        V = so.table("T_re", q) + 1j * so.table("T_im", q)
        Vi = so.table("T_inv_re", q) + 1j * so.table("T_inv_im", q)
        lam = so.table("D_vec_re_", q) + 1j * so.table("D_vec_im_", q)
        assert np.abs(V @ np.diag(lam) @ Vi - Ai).max() < 1e-8 * np.abs(Ai).max()


# (dim,k,r,q) -> step-1 and step-5 (L2, Linf, nodal l2 norm)
KNOWN = {
    (2, 2, 3, 2): ((2.419429e-03, 3.917341e-03, 9.963278140944), (3.006095e-03, 4.760362e-03, 12.46545730316)),
    (3, 4, 2, 4): ((1.116498e-04, 3.658142e-04, 28.17520877213), (1.397713e-04, 4.574229e-04, 35.24462230090)),
    (3, 1, 3, 2): ((2.627479e-02, 1.169853e-01, 11.12709453028), (3.587512e-02, 1.548329e-01, 13.80970241796)),
}


import functools


@functools.lru_cache(maxsize=None)
def direct_run(dim, k, r, q):
    """five direct (sparse LU) steps; returns the per-step solutions."""
    prob = so.Problem(dim, k, r)
    u, t, out = prob.initial(), 0.0, []
    for s in range(5):
        t += 0.1
        u, _ = so.direct_irk_step(prob, q, 0.1, u, t)
        out.append((t, u, prob.errors(u, t)))
    return out


@pytest.mark.parametrize("case", sorted(KNOWN))
def test_direct_solve_known_answers(case):
    steps = direct_run(*case)
    for s in (0, 4):
        ref = KNOWN[case][0 if s == 0 else 1]
        t, u, e = steps[s]
        assert abs(e[0] - ref[0]) < 2e-6 * ref[0] and abs(e[1] - ref[1]) < 2e-6 * ref[1]
        assert abs(np.sqrt(so.dot(u, u)) - ref[2]) < 1e-11 * ref[2]


@pytest.mark.parametrize("scheme,dim,k,r,q,tol,counts", [
    ("irk", 3, 4, 2, 4, 1e-8, [7, 7, 7, 6, 7]),
    ("irk", 3, 4, 2, 4, 1e-12, [10, 10, 10, 9, 10]),
    ("irk", 2, 2, 5, 2, 1e-8, [4, 4, 4, 4, 4]),
])
def test_iterative_matches_direct_and_predicted_counts(scheme, dim, k, r, q, tol, counts):
    o = so.run(scheme, dim, k, r, q, 0.1, 0.5, outer_tol=tol)
    assert o["integ"].n_outer == counts
    u = direct_run(dim, k, r, q)[4][1]
    assert ac.relerr(o["u"], u) < 50 * tol


def test_cpu_port_kernels_match_numpy_oracle(cpu_dev):
    assert cpu_dev.backend() == "cpu-oracle"
    ac.run_all(cpu_dev)


def test_product_library_exports_every_symbol():
    """include/spirk_b200.h <-> libspirk_b200.so: load (no GPU needed) and resolve every symbol."""
    import dealii_spirk_b200
    path = dealii_spirk_b200.DEVICE_LIB_PATH
    if not os.path.exists(path):
        import __graft_entry__
        __graft_entry__.build()
    lib = C.CDLL(path)
    header = open(os.path.join(ROOT, "include", "spirk_b200.h")).read()
    import re
    declared = sorted(set(re.findall(r"\b(spirk_[a-z0-9_]+)\s*\(", header)))
    assert declared == capi.ALL_SYMBOLS, set(declared) ^ set(capi.ALL_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    dev = dealii_spirk_b200.device_lib()
    assert dev.backend() == "cuda-sm_100a"
    import torch
    if not torch.cuda.is_available():
        h = C.c_void_p()
        st = dev.lib.spirk_ctx_create(C.byref(h), 0)
        assert st == 2, "product library must fail loudly (SPIRK_ERR_DEVICE) without a GPU"


def test_host_library_exports_every_symbol():
    """include/spirk_host.h <-> libspirk_host.so (the C++ host layer linked against the CUDA library)."""
    import re
    import dealii_spirk_b200
    path = dealii_spirk_b200.HOST_LIB_PATH
    if not os.path.exists(path):
        import __graft_entry__
        __graft_entry__.build()
    C.CDLL(dealii_spirk_b200.DEVICE_LIB_PATH, mode=C.RTLD_GLOBAL)
    lib = C.CDLL(path)
    header = open(os.path.join(ROOT, "include", "spirk_host.h")).read()
    declared = sorted(set(re.findall(r"\b(spirk_host_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), name


def test_cpu_double_virtual_rank_mixing(cpu_dev):
    """the same check the -m gpu suite runs on the peer mixing kernels, here on the CPU double (host-side logic of the test)"""
    import abi_checks as ac
    ac.check_mix_peer_virtual(cpu_dev, 2, 1, n=2003)
    ac.check_mix_peer_virtual(cpu_dev, 4, 2, n=1001)


@pytest.mark.parametrize("scheme,q,max_diff,min_diff", [("irk", 2, 1e-7, 1e-11), ("spirk", 4, 1e-5, 1e-9), ("spirk", 8, 1e-2, 1e-6)])
def test_fp32_vcycle_study(scheme, q, max_diff, min_diff):
    """SURVEY 8f rank 4 (FP32 V-cycle under the FP64 outer GMRES; reference preconditioner.h:120-142 anticipates a float
    level vector), studied in the oracle because the product has no FP32 kernels: with the reference's solver set-up
    (left-preconditioned, non-flexible GMRES) the outer iteration counts stay within +1 for q <= 4, but the solution moves
    by 1e-8 .. 1e-7 relative (q = 2, 4) and 1e-4 (q = 8, cond(T) = 7e5) - outside the 1e-10 parity bar of the north star.
    DESIGN.md section 7 quotes these numbers as the reason the FP32 V-cycle is not built."""
    a = so.run(scheme, 3, 4, 2, q, 0.1, 0.3, outer_tol=1e-12)
    b = so.run(scheme, 3, 4, 2, q, 0.1, 0.3, outer_tol=1e-12, fp32_vcycle=True)
    d = np.max(np.abs(a["u"] - b["u"])) / np.max(np.abs(a["u"]))
    assert min_diff < d < max_diff, d
    na, nb = np.array(a["integ"].n_outer), np.array(b["integ"].n_outer)
    assert np.all(nb >= na) and np.all(nb - na <= (1 if q <= 4 else 4)), (na, nb)
