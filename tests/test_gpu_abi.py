"""-m gpu: every kernel of the CUDA library, called through the C ABI, against the NumPy oracle."""
import pytest

import abi_checks as ac

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", range(len(ac.OP_CASES)))
def test_op_apply_3d_q4(gpu_dev, case):
    ac.check_op_apply(gpu_dev, 3, 4, 2, ac.OP_CASES[case])


@pytest.mark.parametrize("dim,k,r", [(2, 1, 3), (2, 2, 3), (2, 3, 2), (2, 4, 2), (2, 5, 2), (2, 6, 1),
                                     (3, 1, 3), (3, 2, 2), (3, 3, 2), (3, 5, 1), (3, 6, 1), (3, 4, 0), (3, 4, 1),
                                     (3, 4, 3)])
def test_op_apply_shapes(gpu_dev, dim, k, r):
    ac.check_op_apply(gpu_dev, dim, k, r, ac.OP_CASES[0])
    ac.check_op_apply(gpu_dev, dim, k, r, ac.OP_CASES[3])
    ac.check_op_apply(gpu_dev, dim, k, r, ac.OP_CASES[4])


@pytest.mark.parametrize("dim,k,r,nb", [(3, 4, 2, 2), (3, 4, 1, 4), (2, 2, 3, 1), (3, 1, 3, 2), (3, 4, 3, 2), (3, 4, 4, 1),
                                        (3, 4, 3, 4)])
def test_residual_and_chebyshev_step(gpu_dev, dim, k, r, nb):
    ac.check_residual_and_cheb(gpu_dev, dim, k, r, nb)


@pytest.mark.parametrize("dim,k,r", [(3, 4, 2), (2, 2, 4), (3, 1, 3), (3, 2, 2), (2, 4, 2)])
def test_diag_transfer_problem(gpu_dev, dim, k, r):
    ac.check_inverse_diagonal(gpu_dev, dim, k, r)
    ac.check_transfer(gpu_dev, dim, k, r)
    ac.check_problem(gpu_dev, dim, k, r)


@pytest.mark.parametrize("r,nb", [(3, 2), (4, 1), (5, 2), (4, 3)])
def test_fast_path_matches_general_kernel(gpu_dev, r, nb):
    """variant 0 (plane streaming, op_v3) and variant 2 (pipelined tile columns, op_v2) against
    variant 1 (general cell kernel) and the oracle"""
    import ctypes as C
    import numpy as np
    from dealii_spirk_b200 import capi
    lvl, olv = ac.make_level(3, 4, r)
    u = ac.block_input(olv, nb, seed=11)
    mass = [16.0, 2.9418686642961562, 5.644106850167844][:nb]
    op = capi.real_op(mass, [0.1])
    outs = {}
    with capi.Context(gpu_dev) as ctx:
        src, dst = ctx.upload(u), ctx.alloc(u.size)
        for variant in (1, 2, 0):
            ctx.call("spirk_ctx_set_option", b"apply_variant", variant)
            ctx.call("spirk_op_apply", C.byref(lvl), C.byref(op), dst, src, olv.N)
            outs[variant] = ctx.download(dst, u.shape)
            if variant != 1:
                ctx.call("spirk_op_apply", C.byref(lvl), C.byref(op), dst, src, olv.N)
                again = ctx.download(dst, u.shape)
                assert np.array_equal(outs[variant], again), "fast paths must be bitwise reproducible (no atomics)"
    assert ac.relerr(outs[2], outs[1]) < 1e-13
    assert ac.relerr(outs[0], outs[1]) < 1e-13
    if r <= 4:
        assert ac.relerr(outs[0], olv.apply(u, mass, [0.1] * nb)) < 1e-12


@pytest.mark.parametrize("variant", [0, 2])
@pytest.mark.parametrize("r,nb", [(3, 1), (4, 2)])
def test_fast_path_fused_epilogues(gpu_dev, variant, r, nb):
    """residual / Chebyshev step (explicit and on-the-fly inverse diagonal, aliasing) on both fast paths"""
    from dealii_spirk_b200 import capi

    # set the variant option on every context the check opens
    orig = capi.Context.__enter__

    def enter(self):
        ctx = orig(self)
        ctx.call("spirk_ctx_set_option", b"apply_variant", variant)
        return ctx

    capi.Context.__enter__ = enter
    try:
        ac.check_residual_and_cheb(gpu_dev, 3, 4, r, nb)
    finally:
        capi.Context.__enter__ = orig


def test_assemble_dense(gpu_dev):
    ac.check_assemble_dense(gpu_dev, 3, 4)
    ac.check_assemble_dense(gpu_dev, 2, 2)


def test_vector_kernels(gpu_dev):
    ac.check_vector_ops(gpu_dev)
    ac.check_mgs(gpu_dev)
    ac.check_mix(gpu_dev)
    ac.check_mix(gpu_dev, q=8)
    ac.check_dense_matvec(gpu_dev)


# every k_v3<K, TX, TY, MODE, NPT, NBC> instantiation v3_launch can select, at the sizes the bench runs (r = 5, 6): the
# 8 x 8-cell tiles with NPT = 4 (apply) and NPT = 2 (fused epilogues, and apply under v3_npt = 2), nb in {1, 2}, even-split
# and z-lockstep schedules; the 4 x 4-cell instantiations are the r <= 4 cases above
@pytest.mark.parametrize("r,nb,opts", [
    (5, 1, {"v3_schedule": 0}), (5, 1, {"v3_schedule": 1}), (5, 2, {"v3_schedule": 0, "v3_npt": 2}),
    (5, 2, {"v3_schedule": 1, "v3_npt": 4}), (5, 2, {}), (6, 2, {}), (6, 1, {"v3_schedule": 0}), (6, 2, {"v3_schedule": 1, "v3_npt": 2}),
    (5, 2, {"v3_small_below": 64}),  # 4 x 4-cell tiles at r = 5
    (6, 2, {"v3_tail": 0}),  # equal ranges (the default at this size: long ranges first, short ones for the tail)
])
def test_v3_all_modes_full_size(gpu_dev, r, nb, opts):
    ac.check_v3_full_size(gpu_dev, r, nb, opts)


@pytest.mark.parametrize("r,opts", [(5, {}), (5, {"v3_schedule": 1}), (6, {}), (5, {"v3_small_below": 64})])
def test_v3_coupled_pair_full_size(gpu_dev, r, opts):
    """k_v3<..., APPLY, 4, NBC = 2>: complex pair and the IRK q = 2 system matrix (main.cc:1014-1028) at bench size"""
    ac.check_op_apply(gpu_dev, 3, 4, r, ac.OP_CASES[4], opts=opts)
    ac.check_op_apply(gpu_dev, 3, 4, r, ("coupled", ac.so.table("A_inv", 2).tolist(), [0.1]), opts=opts)


# k_v3<..., MODE, 4, NBC = 2> for MODE = residual / Chebyshev (explicit and own diagonal) / fused first iterations: 4 x 4-cell
# tiles (r = 3, 4 and r = 5 under v3_small_below), 8 x 8-cell tiles (r = 5, 6), work-queue and z-lockstep schedules; unequal
# diagonals (the general pair: the first iterations fall back to scale + step); shapes the fast path does not cover
@pytest.mark.parametrize("dim,k,r,desc,opts", [
    (3, 4, 2, None, {}), (3, 4, 3, None, {}), (3, 4, 4, None, {}), (3, 4, 5, None, {}), (3, 4, 5, None, {"v3_schedule": 1}),
    (3, 4, 5, None, {"v3_small_below": 64}), (3, 4, 6, None, {}),
    (3, 4, 4, ("coupled", [[2.0, 0.5], [-0.25, 3.0]], [0.1, 0.2]), {}), (3, 4, 5, ("coupled", [[2.0, 0.5], [-0.25, 3.0]], [0.1, 0.2]), {}),
    (3, 4, 4, None, {"apply_variant": 1}), (2, 2, 4, None, {}), (3, 2, 2, None, {}),
])
def test_coupled_pair_fused_epilogues(gpu_dev, dim, k, r, desc, opts):
    """residual / smoother kernels of the complex level operators (operator.h:616-665) against the oracle"""
    ac.check_coupled_fused(gpu_dev, dim, k, r, desc, opts)


@pytest.mark.parametrize("dim,k,r,nb,opts", [(3, 4, 2, 2, {}), (3, 4, 3, 1, {}), (3, 4, 4, 3, {}), (3, 4, 5, 2, {}), (3, 4, 5, 1, {"v3_schedule": 1}),
                                             (3, 4, 6, 1, {}), (2, 2, 4, 2, {}), (3, 2, 2, 4, {}), (3, 4, 1, 2, {})])
def test_op_apply_km(gpu_dev, dim, k, r, nb, opts):
    """the one-pass K v + M w operator (k_v3<..., NBC = 2> with the second plane from another vector; general path elsewhere)"""
    ac.check_apply_km(gpu_dev, dim, k, r, nb, opts)


@pytest.mark.parametrize("dim,k,r,nb", [(3, 4, 5, 2), (3, 4, 6, 2), (2, 6, 7, 1), (2, 4, 8, 2)])
def test_transfer_full_size(gpu_dev, dim, k, r, nb):
    """incl. rows of 769 / 1025 nodes: the staged rows of the x sweeps exceed the default dynamic shared-memory limit (as on the
    3-D r = 7 level, 513 nodes per row)"""
    ac.check_transfer(gpu_dev, dim, k, r, nb)


@pytest.mark.parametrize("r,C,nb", [(3, 2, 2), (4, 2, 1), (4, 4, 2), (5, 2, 2), (5, 4, 1), (6, 2, 1)])
def test_slab_kernels(gpu_dev, r, C, nb):
    """z-slab levels (spatial partition): every kernel on every slab against the slice of the whole-mesh oracle result"""
    ac.check_slab_kernels(gpu_dev, r, C, nb)
