"""CPU suite, world_size 2 over gloo: the N>1 host path (stage / conjugate-pair ownership per rank,
all-gather stage mixing, row-reduced Krylov scalars, solution all-reduce; reference main.cc:1229-1760,
2382-2934) must reproduce the single-process oracle: solution 1e-10, iteration counts +-1."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import spirk_oracle as so

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("scheme,dim,k,r,q", [("spirk", 2, 2, 3, 2), ("spirk", 3, 4, 1, 4), ("complex_spirk", 2, 2, 3, 4),
                                              ("complex_spirk_batched", 2, 2, 3, 3)])
def test_two_ranks_match_oracle(tmp_path, scheme, dim, k, r, q):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"])
    out = str(tmp_path / "res.json")
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + (os.getpid() + q) % 300), os.path.join(ROOT, "tests", "gloo_worker.py"), scheme, str(dim),
           str(k), str(r), str(q), out]
    subprocess.run(cmd, check=True, env=env, timeout=600, capture_output=True)
    res = json.load(open(out))
    ora = so.run(scheme, dim, k, r, q, 0.1, 0.5, outer_tol=1e-12)
    uo = ora["u"].reshape(-1)
    assert np.max(np.abs(np.array(res["u"]) - uo)) < 1e-10 * np.max(np.abs(uo))
    n_outer = ora["integ"].n_outer
    ref = np.array([max(o) if isinstance(o, list) else o for o in n_outer])
    assert np.all(np.abs(np.array(res["outer"]) - ref) <= 1), (res["outer"], n_outer)


def test_grid_mapping():
    """lex_to_pair of the reference (main.cc:281-293): stage index = rank % size_x (row major)"""
    from dealii_spirk_b200.launch import lex_to_pair
    assert [lex_to_pair(r, 4, 2, True) for r in range(8)] == [(0, 0), (1, 0), (2, 0), (3, 0), (0, 1), (1, 1), (2, 1), (3, 1)]
    assert [lex_to_pair(r, 4, 2, False) for r in range(8)] == [(0, 0), (0, 1), (1, 0), (1, 1), (2, 0), (2, 1), (3, 0), (3, 1)]
    with pytest.raises(ValueError):
        lex_to_pair(8, 4, 2, True)
