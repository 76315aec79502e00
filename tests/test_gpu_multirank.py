"""-m gpu: the multi-rank CUDA path.

1. The peer-memory mixing kernels (k_mix_peer: fused all-gather + mixing, k_mix_a2a + k_mix_finish: fused all-to-all +
   mixing; reference matrix_vector_rol_operation / perform_basis_change, main.cc:1443-1534) on ONE GPU with the ranks of
   the stage group emulated by a same-device exchange group (B200_PROFILING.md: with fewer GPUs than ranks run all ranks'
   data through the kernels on one device) - always runs.
2. Stage-parallel runs with one process per GPU under torchrun (NCCL, CUDA-IPC exchange buffers, all three SPIRK_PEER_MIX
   modes) against the NumPy oracle: solution 1e-10, iteration counts +-1 - runs when the box has >= 2 (>= 4) GPUs.
"""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import abi_checks as ac
import spirk_oracle as so
from dealii_spirk_b200 import capi

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("R,m", [(2, 1), (4, 1), (8, 1), (2, 2), (4, 2), (3, 1)])
def test_peer_mixing_kernels_virtual_ranks(gpu_dev, R, m):
    ac.check_mix_peer_virtual(gpu_dev, R, m)


def n_gpus():
    import torch
    return torch.cuda.device_count()


def run_ranks(tmp_path, world, scheme, dim, k, r, q, tol=1e-12, env_extra=None):
    out = str(tmp_path / "res.json")
    env = dict(os.environ, **(env_extra or {}))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + (os.getpid() + q + world) % 200), os.path.join(ROOT, "tests", "nccl_worker.py"), scheme,
           str(dim), str(k), str(r), str(q), str(tol), out]
    p = subprocess.run(cmd, env=env, timeout=900, capture_output=True, text=True)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    return json.load(open(out))


def compare_with_oracle(res, scheme, dim, k, r, q, tol=1e-12, sol_tol=1e-10):
    ora = so.run(scheme, dim, k, r, q, 0.1, 0.5, outer_tol=tol)
    uo = ora["u"].reshape(-1)
    rel = np.max(np.abs(np.array(res["u"]) - uo)) / np.max(np.abs(uo))
    assert rel < sol_tol, f"{scheme}: solution differs from the oracle by {rel}"
    assert np.allclose(res["error_L2"], np.array(ora["errors"])[:, 0], rtol=1e-7, atol=0)
    n_outer = ora["integ"].n_outer
    ref = np.array([max(o) if isinstance(o, (list, tuple)) else o for o in n_outer])
    assert np.all(np.abs(np.array(res["outer"]) - ref) <= 1), (res["outer"], n_outer)


# SPIRK_PEER_MIX: 0 = NCCL all-gather + local mixing kernel, 1 = gather kernel over peer memory, unset = all-to-all kernel
@pytest.mark.parametrize("peer_mix", ["0", "1", None])
@pytest.mark.parametrize("scheme,dim,k,r,q", [("spirk", 3, 4, 2, 2), ("spirk", 3, 4, 2, 4), ("complex_spirk_batched", 3, 4, 2, 4),
                                              ("complex_spirk", 2, 2, 3, 3)])
def test_two_gpu_ranks_match_oracle(tmp_path, scheme, dim, k, r, q, peer_mix):
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    env = {} if peer_mix is None else {"SPIRK_PEER_MIX": peer_mix}
    res = run_ranks(tmp_path, 2, scheme, dim, k, r, q, env_extra=env)
    compare_with_oracle(res, scheme, dim, k, r, q)


@pytest.mark.parametrize("scheme,dim,k,r,q", [("spirk", 3, 4, 2, 4), ("spirk", 3, 4, 2, 8), ("complex_spirk_batched", 3, 4, 2, 8)])
def test_four_gpu_ranks_match_oracle(tmp_path, scheme, dim, k, r, q):
    if n_gpus() < 4:
        pytest.skip("needs 4 GPUs (run with gpurun --gpus 4)")
    res = run_ranks(tmp_path, 4, scheme, dim, k, r, q)
    compare_with_oracle(res, scheme, dim, k, r, q, sol_tol=1e-10 if q < 8 else 1e-9)


def test_spirk_ranks_equal_single_gpu_at_baseline_size(tmp_path, gpu_dev):
    """BASELINE size (3-D Q4, r = 5, 2.1e6 DoFs x q = 2 stages): spirk on 2 GPUs against irk on 1 GPU - the same algebra
    (main.cc:815-974 vs 1281-1440): error norms and solution norm 1e-10, iteration counts identical"""
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import host_checks as hc
    import dealii_spirk_b200 as pkg
    from dealii_spirk_b200 import hostapi
    two = run_ranks(tmp_path, 2, "spirk", 3, 4, 5, 2, tol=1e-10)
    host = hostapi.HostLib(pkg.HOST_LIB_PATH, hc.TABLES)
    one = hc.run_host(host, "irk", 3, 4, 5, 2, tol=1e-10)
    assert np.allclose(two["norm"], one["norm"], rtol=1e-10, atol=0)
    assert np.allclose(two["error_L2"], one["error_L2"], rtol=1e-6, atol=0)
    assert np.all(np.abs(np.array(two["outer"]) - one["outer"]) <= 1), (two["outer"], one["outer"])


# ---- spatial partition (z-slabs over the column communicator, halo exchange, replicated coarse levels): SURVEY 8e
# SPIRK_SLAB_MIN_CELLS = 8: every level down to 8 cells per direction is split (slab -> slab transfers with halo exchange on
# both levels); default 32: the levels below the finest are replicated (slab -> replicated transfers with the all-gather)
@pytest.mark.parametrize("scheme,k,r,q,min_cells", [("irk", 4, 3, 2, "8"), ("irk", 4, 4, 2, "8"), ("irk", 4, 4, 2, None), ("ost", 4, 3, 0, None)])
def test_two_slabs_match_oracle(tmp_path, scheme, k, r, q, min_cells):
    """1 stage rank x 2 space ranks: the mesh hierarchy is split into two z-slabs"""
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    res = run_ranks(tmp_path, 2, scheme, 3, k, r, q, env_extra={"SPIRK_SLAB_MIN_CELLS": min_cells} if min_cells else None)
    if scheme == "ost":
        ora = so.run(scheme, 3, k, r, q, 0.1, 0.5, outer_tol=1e-12)
        uo = ora["u"].reshape(-1)
        assert np.max(np.abs(np.array(res["u"]) - uo)) < 1e-8 * np.max(np.abs(uo))  # CG to 1e-8 |rhs| (main.cc:526)
    else:
        compare_with_oracle(res, scheme, 3, k, r, q)


@pytest.mark.parametrize("r,q,min_cells", [(3, 2, None), (4, 2, "8")])
def test_stage_x_space_grid_four_gpus(tmp_path, r, q, min_cells):
    """spirk with q = 2 stage ranks x 2 space ranks on 4 GPUs (the reference's rectangular grid, main.cc:3660-3698)"""
    if n_gpus() < 4:
        pytest.skip("needs 4 GPUs (run with gpurun --gpus 4)")
    res = run_ranks(tmp_path, 4, "spirk", 3, 4, r, q, env_extra={"SPIRK_SLAB_MIN_CELLS": min_cells} if min_cells else None)
    compare_with_oracle(res, "spirk", 3, 4, r, q)
