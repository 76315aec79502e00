"""Checks of any implementation of include/spirk_b200.h against the NumPy oracle.

The same functions run in the CPU suite against the CPU double (oracle/_build/libspirk_cpu.so)
and in the `-m gpu` suite against the product CUDA library through the C ABI.
Tolerances: FP64, relative 1e-12 per kernel application (north_star asks 1e-10 on the solution).
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import spirk_oracle as so  # noqa: E402
from dealii_spirk_b200 import capi  # noqa: E402

RTOL = 1e-12


def synth(n, seed=0):
    """stateless synthetic vector of SURVEY 8(d): 2*frac(sin(12.9898*(i+1))*43758.5453)-1."""
    i = np.arange(n, dtype=np.float64) + 1.0 + 1000.0 * seed
    v = np.sin(12.9898 * i) * 43758.5453
    return 2.0 * (v - np.floor(v)) - 1.0


def relerr(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def make_level(dim, k, r):
    return capi.Level(dim, k, 2 ** r, 0), so.Level(dim, k, r)


def block_input(olv, nb, seed=0, zero_boundary=False):
    u = synth(nb * olv.N, seed).reshape((nb,) + olv.shape)
    if zero_boundary:
        u[:, olv.bmask] = 0.0
    return u


def set_options(ctx, opts):
    """spirk_ctx_set_option knobs (include/spirk_b200.h) for A/B tests of the kernel variants / schedules"""
    for name, value in (opts or {}).items():
        ctx.call("spirk_ctx_set_option", name.encode(), int(value))


def check_op_apply(dev, dim, k, r, desc, tol=RTOL, opts=None):
    """desc: ('real', mass[], lap[]) or ('coupled', C[][], lap[])."""
    lvl, olv = make_level(dim, k, r)
    if desc[0] == "real":
        op = capi.real_op(desc[1], desc[2])
        mass = np.atleast_1d(np.asarray(desc[1], float))
        nb = len(mass)
        lap = np.broadcast_to(np.atleast_1d(np.asarray(desc[2], float)), (nb,))
        u = block_input(olv, nb, seed=1)
        ref = olv.apply(u, mass, lap)
    else:
        Cm = np.asarray(desc[1], float)
        nb = Cm.shape[0]
        lap = np.broadcast_to(np.atleast_1d(np.asarray(desc[2], float)), (nb,))
        op = capi.coupled_op(Cm, lap)
        u = block_input(olv, nb, seed=2)
        v = u.copy()
        v[:, olv.bmask] = 0.0
        Mv = olv.apply(v, 1.0, 0.0)
        Kv = olv.apply(v, 0.0, lap)
        ref = Kv + np.tensordot(Cm, Mv, axes=(1, 0))
        ref[:, olv.bmask] = u[:, olv.bmask]
    with capi.Context(dev) as ctx:
        set_options(ctx, opts)
        src = ctx.upload(u)
        dst = ctx.alloc(u.size)
        ctx.call("spirk_op_apply", C.byref(lvl), C.byref(op), dst, src, olv.N)
        out = ctx.download(dst, u.shape)
    e = relerr(out, ref)
    assert e < tol, f"op_apply {desc[0]} dim={dim} k={k} r={r}: rel err {e}"
    return e


def check_apply_km(dev, dim, k, r, nb=2, opts=None):
    """dst_b = laplace_b K v_b + mass_b M w_b in one cell pass (the stage-parallel system matrix after the A_inv mixing of
    the source, reference main.cc:1580-1592); Dirichlet rows dst = v"""
    lvl, olv = make_level(dim, k, r)
    v, w = block_input(olv, nb, 21), block_input(olv, nb, 22)
    lap = np.array([0.1, 0.2, 0.05, 0.3][:nb])
    mass = np.array([1.0, 2.5, 0.7, 1.3][:nb])
    vz, wz = v.copy(), w.copy()
    vz[:, olv.bmask] = 0.0
    wz[:, olv.bmask] = 0.0
    ref = olv.apply(vz, 0.0, lap) + olv.apply(wz, mass, 0.0)
    ref[:, olv.bmask] = v[:, olv.bmask]
    with capi.Context(dev) as ctx:
        set_options(ctx, opts)
        dv, dw, dd = ctx.upload(v), ctx.upload(w), ctx.alloc(v.size)
        pl, _1 = capi.darr(lap)
        pm, _2 = capi.darr(mass)
        ctx.call("spirk_op_apply_km", C.byref(lvl), nb, dd, dv, dw, olv.N, pl, pm)
        out = ctx.download(dd, v.shape)
    e = relerr(out, ref)
    assert e < RTOL, f"op_apply_km dim={dim} k={k} r={r} nb={nb}: rel err {e}"


def check_residual_and_cheb(dev, dim, k, r, nb=2):
    lvl, olv = make_level(dim, k, r)
    mass = np.array([16.0, 3.1618475338398158, 2.9418686642961562, 5.644106850167844][:nb])
    lap = np.full(nb, 0.1)
    op = capi.real_op(mass, lap)
    x = block_input(olv, nb, 3, True)
    xo = block_input(olv, nb, 4, True)
    b = block_input(olv, nb, 5, True)
    dinv = np.concatenate([olv.inverse_diagonal(m, 0.1) for m in mass])
    Ax = olv.apply(x, mass, lap)
    f1 = np.array([0.3, 0.2, 0.25, 0.35][:nb])
    f2 = np.array([1.1, 0.9, 1.0, 1.2][:nb])
    bc = (slice(None),) + (None,) * dim
    with capi.Context(dev) as ctx:
        dx, dxo, db, dd = (ctx.upload(a) for a in (x, xo, b, dinv))
        dres, dnew = ctx.alloc(x.size), ctx.alloc(x.size)
        ctx.call("spirk_op_residual", C.byref(lvl), C.byref(op), dres, db, dx, olv.N)
        res = ctx.download(dres, x.shape)
        assert relerr(res, b - Ax) < RTOL
        pf1, _a = capi.darr(f1)
        pf2, _b = capi.darr(f2)
        ctx.call("spirk_op_cheb_step", C.byref(lvl), C.byref(op), dnew, dx, dxo, db, dd, olv.N, pf1, pf2)
        new = ctx.download(dnew, x.shape)
        ref = x + f1[bc] * (x - xo) + f2[bc] * dinv * (b - Ax)
        assert relerr(new, ref) < RTOL
        ctx.call("spirk_op_cheb_step", C.byref(lvl), C.byref(op), dnew, dx, None, db, dd, olv.N, pf1, pf2)
        new = ctx.download(dnew, x.shape)
        ref = (1 + f1[bc]) * x + f2[bc] * dinv * (b - Ax)
        assert relerr(new, ref) < RTOL
        # dinv == NULL: the operator's own inverse diagonal, computed on the fly
        ctx.call("spirk_op_cheb_step", C.byref(lvl), C.byref(op), dnew, dx, dxo, db, None, olv.N, pf1, pf2)
        new = ctx.download(dnew, x.shape)
        ref = x + f1[bc] * (x - xo) + f2[bc] * dinv * (b - Ax)
        assert relerr(new, ref) < RTOL
        # iterations 0 and 1 from a zero start in one pass (spirk_op_cheb_first)
        f0 = np.array([0.7, 0.6, 0.65, 0.75][:nb])
        pf0, _c = capi.darr(f0)
        dx1, dx2 = ctx.alloc(x.size), ctx.alloc(x.size)
        bb = block_input(olv, nb, 6, False)  # boundary values of the right-hand side take part (identity rows)
        dbb = ctx.upload(bb)
        ctx.call("spirk_op_cheb_first", C.byref(lvl), C.byref(op), dx1, dx2, dbb, olv.N, pf0, pf1, pf2)
        x1 = f0[bc] * dinv * bb
        x2 = x1 + f1[bc] * x1 + f2[bc] * dinv * (bb - olv.apply(x1, mass, lap))
        assert relerr(ctx.download(dx1, x.shape), x1) < RTOL
        assert relerr(ctx.download(dx2, x.shape), x2) < RTOL
        # x_new aliasing x_old (how the smoother calls it)
        ctx.call("spirk_op_cheb_step", C.byref(lvl), C.byref(op), dxo, dx, dxo, db, None, olv.N, pf1, pf2)
        assert relerr(ctx.download(dxo, x.shape), ref) < RTOL



def check_coupled_fused(dev, dim, k, r, desc=None, opts=None, tol=RTOL):
    """Residual and Chebyshev steps of a COUPLED pair of blocks (the complex level operator, reference operator.h:616-665,
    under the smoother of preconditioner.h:353-373): explicit inverse diagonal, the operator's own diagonal (dinv == NULL:
    the diagonal of block b's own term coupling[b][b] M + laplace[b] K), x_old = 0, x_new aliasing x_old, and the fused
    first two iterations; against the Kronecker oracle."""
    desc = desc or OP_CASES[4]
    lvl, olv = make_level(dim, k, r)
    Cm = np.asarray(desc[1], float)
    nb = Cm.shape[0]
    lap = np.broadcast_to(np.atleast_1d(np.asarray(desc[2], float)), (nb,)).copy()
    op = capi.coupled_op(Cm, lap)

    def apply(u):
        v = u.copy()
        v[:, olv.bmask] = 0.0
        out = olv.apply(v, 0.0, lap) + np.tensordot(Cm, olv.apply(v, 1.0, 0.0), axes=(1, 0))
        out[:, olv.bmask] = u[:, olv.bmask]
        return out

    x, xo, b = block_input(olv, nb, 3, True), block_input(olv, nb, 4, True), block_input(olv, nb, 5, True)
    bb = block_input(olv, nb, 6, False)
    own = np.concatenate([olv.inverse_diagonal(Cm[i, i], lap[i]) for i in range(nb)])
    other = np.concatenate([olv.inverse_diagonal(1.0 + i, 1.0) for i in range(nb)])  # a diagonal that is NOT the operator's
    f0, f1, f2 = np.array([0.7, 0.6][:nb]), np.array([0.3, 0.2][:nb]), np.array([1.1, 0.9][:nb])
    bc = (slice(None),) + (None,) * dim
    Ax = apply(x)
    errs = {}
    with capi.Context(dev) as ctx:
        set_options(ctx, opts)
        dx, dxo, db, dbb, dd = (ctx.upload(a) for a in (x, xo, b, bb, other))
        d1, d2 = ctx.alloc(x.size), ctx.alloc(x.size)
        pf0, _0 = capi.darr(f0)
        pf1, _1 = capi.darr(f1)
        pf2, _2 = capi.darr(f2)
        N, shape = olv.N, x.shape
        ctx.call("spirk_op_residual", C.byref(lvl), C.byref(op), d1, db, dx, N)
        errs["residual"] = relerr(ctx.download(d1, shape), b - Ax)
        ctx.call("spirk_op_cheb_step", C.byref(lvl), C.byref(op), d1, dx, dxo, db, dd, N, pf1, pf2)
        errs["cheb_dinv"] = relerr(ctx.download(d1, shape), x + f1[bc] * (x - xo) + f2[bc] * other.reshape(shape) * (b - Ax))
        ref = x + f1[bc] * (x - xo) + f2[bc] * own.reshape(shape) * (b - Ax)
        ctx.call("spirk_op_cheb_step", C.byref(lvl), C.byref(op), d1, dx, dxo, db, None, N, pf1, pf2)
        errs["cheb_own"] = relerr(ctx.download(d1, shape), ref)
        if dim == 3 and k == 4 and r >= 3 and not (opts or {}).get("apply_variant"):  # (the general cell kernel uses atomics)
            ctx.call("spirk_op_cheb_step", C.byref(lvl), C.byref(op), d2, dx, dxo, db, None, N, pf1, pf2)
            assert np.array_equal(ctx.download(d1, shape), ctx.download(d2, shape)), "bitwise reproducible"
        ctx.call("spirk_op_cheb_step", C.byref(lvl), C.byref(op), d1, dx, None, db, None, N, pf1, pf2)
        errs["cheb_own_x0"] = relerr(ctx.download(d1, shape), (1 + f1[bc]) * x + f2[bc] * own.reshape(shape) * (b - Ax))
        ctx.call("spirk_op_cheb_first", C.byref(lvl), C.byref(op), d1, d2, dbb, N, pf0, pf1, pf2)
        x1 = f0[bc] * own.reshape(shape) * bb
        x2 = x1 + f1[bc] * x1 + f2[bc] * own.reshape(shape) * (bb - apply(x1))
        errs["cheb_first_x1"] = relerr(ctx.download(d1, shape), x1)
        errs["cheb_first_x2"] = relerr(ctx.download(d2, shape), x2)
        # the diagonal of OTHER coefficients, formed on the fly (a smoother set up before the operator's coefficients changed)
        pdm, _3 = capi.darr([1.0 + i for i in range(nb)])
        pdl, _4 = capi.darr([1.0] * nb)
        ctx.call("spirk_op_cheb_step_diag", C.byref(lvl), C.byref(op), d1, dx, dxo, db, pdm, pdl, N, pf1, pf2)
        errs["cheb_diag"] = relerr(ctx.download(d1, shape), x + f1[bc] * (x - xo) + f2[bc] * other.reshape(shape) * (b - Ax))
        o2 = np.concatenate([olv.inverse_diagonal(1.5, 0.7)] * nb).reshape(shape)
        pdm2, _5 = capi.darr([1.5] * nb)
        pdl2, _6 = capi.darr([0.7] * nb)
        ctx.call("spirk_op_cheb_first_diag", C.byref(lvl), C.byref(op), d1, d2, dbb, pdm2, pdl2, N, pf0, pf1, pf2)
        y1 = f0[bc] * o2 * bb
        errs["cheb_first_diag_x1"] = relerr(ctx.download(d1, shape), y1)
        errs["cheb_first_diag_x2"] = relerr(ctx.download(d2, shape), y1 + f1[bc] * y1 + f2[bc] * o2 * (bb - apply(y1)))
        ctx.call("spirk_op_cheb_step", C.byref(lvl), C.byref(op), dxo, dx, dxo, db, None, N, pf1, pf2)  # x_new aliases x_old
        errs["cheb_own_alias"] = relerr(ctx.download(dxo, shape), ref)
    bad = {k_: v for k_, v in errs.items() if not v < tol}
    assert not bad, f"coupled fused dim={dim} k={k} r={r} opts={opts}: {bad}"
    return errs


_FULL_CACHE = {}


def full_size_case(r, nb):
    """oracle data of the 3-D Q4 fused-kernel checks at refinement r (cached: the Kronecker oracle costs seconds at r = 6)"""
    key = (r, nb)
    if key not in _FULL_CACHE:
        _, olv = make_level(3, 4, r)
        mass = np.array(D4[:nb])
        lap = np.full(nb, 0.1)
        x = block_input(olv, nb, 3, True)
        xo = block_input(olv, nb, 4, True)
        b = block_input(olv, nb, 5, True)
        bb = block_input(olv, nb, 6, False)
        dinv = np.concatenate([olv.inverse_diagonal(m, 0.1) for m in mass])
        f0, f1, f2 = np.array([0.7, 0.6, 0.65, 0.75][:nb]), np.array([0.3, 0.2, 0.25, 0.35][:nb]), np.array([1.1, 0.9, 1.0, 1.2][:nb])
        bc = (slice(None),) + (None,) * 3
        Ax = olv.apply(x, mass, lap)
        x1 = f0[bc] * dinv * bb
        x2 = x1 + f1[bc] * x1 + f2[bc] * dinv * (bb - olv.apply(x1, mass, lap))
        dm, dl = 1.3 * mass + 0.2, np.full(nb, 0.25)
        dinv2 = np.concatenate([olv.inverse_diagonal(m, 0.25) for m in dm]).reshape(x.shape)
        y1 = f0[bc] * dinv2 * bb
        y2 = y1 + f1[bc] * y1 + f2[bc] * dinv2 * (bb - olv.apply(y1, mass, lap))
        _FULL_CACHE[key] = dict(olv=olv, mass=mass, lap=lap, x=x, xo=xo, b=b, bb=bb, dinv=dinv, f0=f0, f1=f1, f2=f2, Ax=Ax,
                                dm=dm, dl=dl, y1=y1, y2=y2, cheb_diag=x + f1[bc] * (x - xo) + f2[bc] * dinv2 * (b - Ax),
                                res=b - Ax, cheb=x + f1[bc] * (x - xo) + f2[bc] * dinv * (b - Ax),
                                cheb0=(1 + f1[bc]) * x + f2[bc] * dinv * (b - Ax), x1=x1, x2=x2)
    return _FULL_CACHE[key]


def check_v3_full_size(dev, r, nb, opts=None, tol=RTOL):
    """Every mode of the plane-streaming cell operator the solver launches (apply, residual, Chebyshev step with explicit /
    on-the-fly inverse diagonal, x_old = 0, x_new aliasing x_old, the fused first two iterations) at a BASELINE-size level
    (r = 5: 2.1e6, r = 6: 1.7e7 DoFs per block) against the Kronecker oracle, under the given schedule / thread-layout options."""
    c = full_size_case(r, nb)
    olv, N = c["olv"], c["olv"].N
    lvl = capi.Level(3, 4, 2 ** r, 0)
    op = capi.real_op(c["mass"], c["lap"])
    errs = {}
    with capi.Context(dev) as ctx:
        set_options(ctx, opts)
        dx, dxo, db, dd, dbb = (ctx.upload(c[k]) for k in ("x", "xo", "b", "dinv", "bb"))
        d1, d2 = ctx.alloc(nb * N), ctx.alloc(nb * N)
        shape = c["x"].shape
        pf0, _0 = capi.darr(c["f0"])
        pf1, _1 = capi.darr(c["f1"])
        pf2, _2 = capi.darr(c["f2"])
        ctx.call("spirk_op_apply", C.byref(lvl), C.byref(op), d1, dx, N)
        errs["apply"] = relerr(ctx.download(d1, shape), c["Ax"])
        ctx.call("spirk_op_apply", C.byref(lvl), C.byref(op), d2, dx, N)
        assert np.array_equal(ctx.download(d1, shape), ctx.download(d2, shape)), "the fast path must be bitwise reproducible"
        ctx.call("spirk_op_residual", C.byref(lvl), C.byref(op), d1, db, dx, N)
        errs["residual"] = relerr(ctx.download(d1, shape), c["res"])
        ctx.call("spirk_op_cheb_step", C.byref(lvl), C.byref(op), d1, dx, dxo, db, dd, N, pf1, pf2)
        errs["cheb_dinv"] = relerr(ctx.download(d1, shape), c["cheb"])
        ctx.call("spirk_op_cheb_step", C.byref(lvl), C.byref(op), d1, dx, dxo, db, None, N, pf1, pf2)
        errs["cheb_own"] = relerr(ctx.download(d1, shape), c["cheb"])
        ctx.call("spirk_op_cheb_step", C.byref(lvl), C.byref(op), d1, dx, None, db, None, N, pf1, pf2)
        errs["cheb_own_x0"] = relerr(ctx.download(d1, shape), c["cheb0"])
        ctx.call("spirk_op_cheb_first", C.byref(lvl), C.byref(op), d1, d2, dbb, N, pf0, pf1, pf2)
        errs["cheb_first_x1"] = relerr(ctx.download(d1, shape), c["x1"])
        errs["cheb_first_x2"] = relerr(ctx.download(d2, shape), c["x2"])
        # the diagonal of other coefficients than the operator's, formed on the fly (spirk_op_cheb_step_diag / _first_diag)
        pdm, _3 = capi.darr(c["dm"])
        pdl, _4 = capi.darr(c["dl"])
        ctx.call("spirk_op_cheb_step_diag", C.byref(lvl), C.byref(op), d1, dx, dxo, db, pdm, pdl, N, pf1, pf2)
        errs["cheb_diag"] = relerr(ctx.download(d1, shape), c["cheb_diag"])
        ctx.call("spirk_op_cheb_first_diag", C.byref(lvl), C.byref(op), d1, d2, dbb, pdm, pdl, N, pf0, pf1, pf2)
        errs["cheb_first_diag_x1"] = relerr(ctx.download(d1, shape), c["y1"])
        errs["cheb_first_diag_x2"] = relerr(ctx.download(d2, shape), c["y2"])
        ctx.call("spirk_op_cheb_step", C.byref(lvl), C.byref(op), dxo, dx, dxo, db, None, N, pf1, pf2)  # x_new aliases x_old
        errs["cheb_own_alias"] = relerr(ctx.download(dxo, shape), c["cheb"])
    bad = {k: v for k, v in errs.items() if not v < tol}
    assert not bad, f"r={r} nb={nb} opts={opts}: {bad}"
    return errs


# ------------------------------------------------------------------------------------------------------------
# z-slab levels (spatial partition): every kernel on slab `c` of `C` against the slice of the oracle's result on
# the whole mesh.  Runs on ONE device: the slabs are processed one after the other, the ghost planes a halo exchange would
# bring are copied from the full arrays by the test; planes that no neighbour owns (below the first / above the last slab)
# hold NaN, so any read of them poisons the result.
PAD_HI = 1


def slab_geometry(olv, k, c, C):
    nc = olv.nc if hasattr(olv, "nc") else (olv.shape[0] - 1) // k
    n1 = olv.shape[0]
    L_lo, L_hi = nc * c // C, nc * (c + 1) // C
    zo0, zo1 = k * L_lo, (n1 if L_hi == nc else k * L_hi)
    return dict(nc=nc, n1=n1, L_lo=L_lo, L_hi=L_hi, zo0=zo0, zo1=zo1, pad_lo=2 * k, plane=n1 * n1)


def slab_pack(full, gm):
    """(nb, n1, n1, n1) -> per block [pad_lo | owned | pad_hi] planes; returns the flat buffer, the offset of the first owned
    entry and the block stride"""
    nb, n1 = full.shape[0], gm["n1"]
    nz = gm["pad_lo"] + (gm["zo1"] - gm["zo0"]) + PAD_HI
    buf = np.full((nb, nz, n1, n1), np.nan)
    for lz in range(nz):
        z = gm["zo0"] - gm["pad_lo"] + lz
        if 0 <= z < n1:
            buf[:, lz] = full[:, z]
    return buf.reshape(-1), gm["pad_lo"] * gm["plane"], nz * gm["plane"]


def slab_owned(flat, gm, nb):
    nz = gm["pad_lo"] + (gm["zo1"] - gm["zo0"]) + PAD_HI
    return flat.reshape(nb, nz, gm["n1"], gm["n1"])[:, gm["pad_lo"]:gm["pad_lo"] + gm["zo1"] - gm["zo0"]]


class SlabBuf:
    """device copy of a slab-local block vector"""

    def __init__(self, ctx, full, gm):
        flat, self.off, self.stride = slab_pack(full, gm)
        self.ctx, self.gm, self.nb, self.size = ctx, gm, full.shape[0], flat.size
        self.base = ctx.upload(flat)
        self.ptr = C.c_void_p(self.base.value + 8 * self.off)

    def owned(self):
        return slab_owned(self.ctx.download(self.base, (self.size,)), self.gm, self.nb)


def check_slab_kernels(dev, r, C_, nb=2, k=4):
    dim = 3
    _, olv = make_level(dim, k, r)
    _, olc = make_level(dim, k, r - 1)
    n1 = olv.shape[0]
    mass = np.array(D4[:nb])
    lap = np.full(nb, 0.1)
    x, xo, b = block_input(olv, nb, 3, True), block_input(olv, nb, 4, True), block_input(olv, nb, 5, True)
    bb = block_input(olv, nb, 6, False)
    w = block_input(olv, nb, 7, False)
    dinv = np.concatenate([olv.inverse_diagonal(m, 0.1) for m in mass])
    f0, f1, f2 = np.array([0.7, 0.6, 0.65, 0.75][:nb]), np.array([0.3, 0.2, 0.25, 0.35][:nb]), np.array([1.1, 0.9, 1.0, 1.2][:nb])
    bc = (slice(None),) + (None,) * 3
    Ax = olv.apply(x, mass, lap)
    Abb = olv.apply(bb, mass, lap)
    x1 = f0[bc] * dinv * bb
    x2 = x1 + f1[bc] * x1 + f2[bc] * dinv * (bb - olv.apply(x1, mass, lap))
    cheb = x + f1[bc] * (x - xo) + f2[bc] * dinv * (b - Ax)
    bz, wz = bb.copy(), w.copy()
    bz[:, olv.bmask] = 0.0
    wz[:, olv.bmask] = 0.0
    km = olv.apply(bz, 0.0, lap) + olv.apply(wz, mass, 0.0)
    km[:, olv.bmask] = bb[:, olv.bmask]
    Cm = np.array([[5.0, 3.0], [-3.0, 5.0]])
    u2 = block_input(olv, 2, 9, False)
    v2 = u2.copy()
    v2[:, olv.bmask] = 0.0
    cpl = olv.apply(v2, 0.0, [0.1, 0.1]) + np.tensordot(Cm, olv.apply(v2, 1.0, 0.0), axes=(1, 0))
    cpl[:, olv.bmask] = u2[:, olv.bmask]
    g = so.GMG(dim, k, r, lambda lv: so.ScalarOp(lv))
    uc = block_input(olc, nb, 11)
    uf0 = block_input(olv, nb, 12)
    prol = uf0 + g.prolongate(r, uc)
    rest = g.restrict(r, bb)
    prob = so.Problem(dim, k, r)
    u_ex = prob.exact_nodal(0.3)
    err_full = prob.errors(u_ex * 1.01, 0.3)
    op = capi.real_op(mass, lap)
    opc = capi.coupled_op(Cm, [0.1, 0.1])
    pf0, _0 = capi.darr(f0)
    pf1, _1 = capi.darr(f1)
    pf2, _2 = capi.darr(f2)
    pl, _3 = capi.darr(lap)
    pm, _4 = capi.darr(mass)
    l2sq_sum, linf_max, dot_sum = 0.0, 0.0, 0.0
    for c in range(C_):
        gm = slab_geometry(olv, k, c, C_)
        gmc = dict(slab_geometry(olc, k, c, C_))
        sl = (slice(None), slice(gm["zo0"], gm["zo1"]))
        slc = (slice(None), slice(gmc["zo0"], gmc["zo1"]))
        for coarse_replicated in (0, 1):
            lvl = capi.Level(dim, k, 2 ** r, c | (C_ << 8) | (coarse_replicated << 16))
            with capi.Context(dev) as ctx:
                N = lvl_n = int(dev.lib.spirk_level_n_dofs(C.byref(lvl)))
                assert N == (gm["zo1"] - gm["zo0"]) * gm["plane"]
                tag = f"r={r} slab {c}/{C_} nb={nb}"
                if coarse_replicated == 0:
                    X, XO, B, BB, W, DI = (SlabBuf(ctx, a, gm) for a in (x, xo, b, bb, w, dinv))
                    D1, D2 = SlabBuf(ctx, np.zeros_like(x), gm), SlabBuf(ctx, np.zeros_like(x), gm)
                    st = X.stride
                    ctx.call("spirk_op_apply", C.byref(lvl), C.byref(op), D1.ptr, X.ptr, st)
                    got = D1.owned()
                    if not relerr(got, Ax[sl]) < RTOL:  # which planes / rows / columns are off
                        d = np.abs(got - Ax[sl])
                        bad = np.argwhere(~(d < 1e-9 * np.max(np.abs(Ax))))
                        raise AssertionError(f"{tag} apply: {len(bad)} entries off; blocks {sorted(set(bad[:, 0]))}, local planes "
                                             f"{sorted(set(bad[:, 1]))}, rows {sorted(set(bad[:, 2]))[:12]}, columns {sorted(set(bad[:, 3]))[:12]}; "
                                             f"first {bad[0].tolist()}: {got[tuple(bad[0])]} vs {Ax[sl][tuple(bad[0])]}")
                    ctx.call("spirk_op_residual", C.byref(lvl), C.byref(op), D1.ptr, B.ptr, X.ptr, st)
                    assert relerr(D1.owned(), (b - Ax)[sl]) < RTOL, tag + " residual"
                    ctx.call("spirk_op_cheb_step", C.byref(lvl), C.byref(op), D1.ptr, X.ptr, XO.ptr, B.ptr, DI.ptr, st, pf1, pf2)
                    assert relerr(D1.owned(), cheb[sl]) < RTOL, tag + " cheb (explicit diagonal)"
                    ctx.call("spirk_op_cheb_step", C.byref(lvl), C.byref(op), D1.ptr, X.ptr, XO.ptr, B.ptr, None, st, pf1, pf2)
                    assert relerr(D1.owned(), cheb[sl]) < RTOL, tag + " cheb (own diagonal)"
                    ctx.call("spirk_op_cheb_first", C.byref(lvl), C.byref(op), D1.ptr, D2.ptr, BB.ptr, st, pf0, pf1, pf2)
                    assert relerr(D1.owned(), x1[sl]) < RTOL and relerr(D2.owned(), x2[sl]) < RTOL, tag + " cheb_first"
                    ctx.call("spirk_op_apply_km", C.byref(lvl), nb, D1.ptr, BB.ptr, W.ptr, st, pl, pm)
                    assert relerr(D1.owned(), km[sl]) < RTOL, tag + " apply_km"
                    U2, DC = SlabBuf(ctx, u2, gm), SlabBuf(ctx, np.zeros_like(u2), gm)
                    ctx.call("spirk_op_apply", C.byref(lvl), C.byref(opc), DC.ptr, U2.ptr, U2.stride)
                    assert relerr(DC.owned(), cpl[sl]) < RTOL, tag + " coupled pair"
                    # setup / problem kernels on the owned range (single block)
                    ctx.call("spirk_op_inverse_diagonal", C.byref(lvl), D1.ptr, 16.0, 0.1)
                    assert relerr(D1.owned()[:1], olv.inverse_diagonal(16.0, 0.1)[sl]) < RTOL, tag + " inverse diagonal"
                    ctx.call("spirk_problem_rhs_spatial", C.byref(lvl), D1.ptr)
                    assert relerr(D1.owned()[:1], prob.rspace[sl]) < 1e-12, tag + " rhs"
                    ctx.call("spirk_problem_interpolate_solution", C.byref(lvl), D1.ptr, 0.3)
                    assert relerr(D1.owned()[:1], u_ex[sl]) < 1e-12, tag + " interpolation"
                    UE = SlabBuf(ctx, u_ex * 1.01, gm)
                    l2, li = C.c_double(), C.c_double()
                    ctx.call("spirk_problem_error_norms_partial", C.byref(lvl), UE.ptr, 0.3, C.byref(l2), C.byref(li))
                    l2sq_sum += l2.value
                    linf_max = max(linf_max, li.value)
                    ctx.call("spirk_constraints_set_zero", C.byref(lvl), nb, BB.ptr, BB.stride)
                    assert np.array_equal(BB.owned(), bz[sl]), tag + " set_zero"
                    dot_sum += ctx.scalar_call("spirk_vec_dot_strided", X.ptr, B.ptr, N, nb, st)
                # transfers: the coarse vector as a slab of the coarse level, or held in full (agglomerated coarse levels)
                F0, FB = SlabBuf(ctx, uf0, gm), SlabBuf(ctx, bb, gm)
                if coarse_replicated:
                    dc = ctx.upload(uc)
                    ctx.call("spirk_mg_prolongate_add", C.byref(lvl), nb, F0.ptr, F0.stride, dc, olc.N)
                    assert relerr(F0.owned(), prol[sl]) < RTOL, tag + " prolongation from a replicated coarse level"
                    dr = ctx.upload(np.full_like(uc, np.nan))
                    ctx.call("spirk_mg_restrict", C.byref(lvl), nb, dr, olc.N, FB.ptr, FB.stride)
                    out = ctx.download(dr, uc.shape)
                    assert relerr(out[slc], rest[slc]) < RTOL, tag + " restriction into a replicated coarse level"
                    lo, hi = gmc["zo0"], gmc["zo1"]
                    assert np.all(np.isnan(out[:, :lo])) and np.all(np.isnan(out[:, hi:])), tag + " restriction wrote outside its planes"
                else:
                    UC = SlabBuf(ctx, uc, gmc)
                    ctx.call("spirk_mg_prolongate_add", C.byref(lvl), nb, F0.ptr, F0.stride, UC.ptr, UC.stride)
                    assert relerr(F0.owned(), prol[sl]) < RTOL, tag + " prolongation"
                    RC = SlabBuf(ctx, np.full_like(uc, np.nan), gmc)
                    ctx.call("spirk_mg_restrict", C.byref(lvl), nb, RC.ptr, RC.stride, FB.ptr, FB.stride)
                    assert relerr(RC.owned(), rest[slc]) < RTOL, tag + " restriction"
    assert abs(np.sqrt(l2sq_sum) - err_full[0]) < 1e-9 * err_full[0] and abs(linf_max - err_full[1]) < 1e-9 * err_full[1]
    assert abs(dot_sum - float(np.sum(x * b))) < 1e-9 * np.sqrt(x.size)


def check_inverse_diagonal(dev, dim, k, r, mass=16.0, lap=0.1):
    lvl, olv = make_level(dim, k, r)
    with capi.Context(dev) as ctx:
        d = ctx.alloc(olv.N)
        ctx.call("spirk_op_inverse_diagonal", C.byref(lvl), d, mass, lap)
        out = ctx.download(d, (1,) + olv.shape)
    assert relerr(out, olv.inverse_diagonal(mass, lap)) < RTOL


def check_assemble_dense(dev, dim, k):
    lvl, olv = make_level(dim, k, 0)
    N = olv.N
    with capi.Context(dev) as ctx:
        A = np.zeros((N, N))
        ctx.call("spirk_op_assemble_dense", C.byref(lvl), 3.0, 0.1, A.ctypes.data_as(C.c_void_p))
    ref = np.zeros((N, N))
    for j in range(N):
        e = np.zeros(N)
        e[j] = 1
        ref[:, j] = olv.apply(e.reshape((1,) + olv.shape), 3.0, 0.1).reshape(-1)
    assert relerr(A, ref) < RTOL


def check_transfer(dev, dim, k, r, nb=2):
    lvl, olv = make_level(dim, k, r)
    g = so.GMG(dim, k, r, lambda lv: so.ScalarOp(lv))
    olc = g.levels[r - 1]
    uc = block_input(olc, nb, 6)
    uf0 = block_input(olv, nb, 7)
    uf = block_input(olv, nb, 8)
    with capi.Context(dev) as ctx:
        dc, df = ctx.upload(uc), ctx.upload(uf0)
        ctx.call("spirk_mg_prolongate_add", C.byref(lvl), nb, df, olv.N, dc, olc.N)
        out = ctx.download(df, uf0.shape)
        assert relerr(out, uf0 + g.prolongate(r, uc)) < RTOL
        df2, dc2 = ctx.upload(uf), ctx.alloc(uc.size)
        ctx.call("spirk_mg_restrict", C.byref(lvl), nb, dc2, olc.N, df2, olv.N)
        out = ctx.download(dc2, uc.shape)
        assert relerr(out, g.restrict(r, uf)) < RTOL


def check_vector_ops(dev, n=100003):
    x, y, z = synth(n, 1), synth(n, 2), synth(n, 3)
    with capi.Context(dev) as ctx:
        dx, dy, dz = ctx.upload(x), ctx.upload(y), ctx.upload(z)
        assert abs(ctx.scalar_call("spirk_vec_dot", dx, dy, n) - x @ y) < 1e-10 * n ** 0.5
        assert abs(ctx.scalar_call("spirk_vec_sum", dx, n) - x.sum()) < 1e-10 * n ** 0.5
        ctx.call("spirk_vec_axpy", dy, 0.5, dx, n)
        y = y + 0.5 * x
        ctx.call("spirk_vec_sadd", dy, 2.0, -1.5, dz, n)
        y = 2.0 * y - 1.5 * z
        ctx.call("spirk_vec_add2", dy, 0.25, dx, -0.75, dz, n)
        y = y + 0.25 * x - 0.75 * z
        ctx.call("spirk_vec_scale", dy, n, 1.25)
        y = 1.25 * y
        assert relerr(ctx.download(dy, (n,)), y) < 1e-14
        r = ctx.scalar_call("spirk_vec_add_and_dot", dy, -0.3, dx, dz, n)
        y = y - 0.3 * x
        assert abs(r - y @ z) < 1e-10 * n ** 0.5
        assert relerr(ctx.download(dy, (n,)), y) < 1e-14
        ctx.call("spirk_vec_equ", dz, 3.0, dx, n)
        assert relerr(ctx.download(dz, (n,)), 3.0 * x) < 1e-15
        ctx.call("spirk_vec_copy", dz, dy, n)
        assert np.array_equal(ctx.download(dz, (n,)), ctx.download(dy, (n,)))
        ctx.call("spirk_vec_set", dz, n, 0.0)
        assert ctx.scalar_call("spirk_vec_dot", dz, dz, n) == 0.0
        # pointwise
        nb, m = 3, n // 3
        f = np.array([0.5, 2.0, -1.0])
        pf, _ = capi.darr(f)
        ctx.call("spirk_vec_scale_pointwise", nb, m, dz, dx, dy, m, pf)
        ref = (f[:, None] * (x[:nb * m].reshape(nb, m) * y[:nb * m].reshape(nb, m))).reshape(-1)
        assert relerr(ctx.download(dz, (nb * m,)), ref) < 1e-14


def check_mgs(dev, n=50021, dim=5):
    rng = np.random.default_rng(1)
    Q, _ = np.linalg.qr(rng.standard_normal((n, dim)))
    basis = np.ascontiguousarray(Q.T)
    vv = synth(n, 9)
    h = np.zeros(dim)
    w = vv.copy()
    for i in range(dim):
        h[i] = w @ basis[i]
        w = w - h[i] * basis[i]
    with capi.Context(dev) as ctx:
        dv, db = ctx.upload(vv), ctx.upload(basis)
        hh = np.zeros(dim)
        nrm = C.c_double()
        ptrs = (C.c_void_p * dim)(*[db.value + 8 * n * i for i in range(dim)])
        ctx.call("spirk_gmres_mgs", dv, ptrs, dim, n, hh.ctypes.data_as(capi.dp), C.byref(nrm))
        out = ctx.download(dv, (n,))
    assert np.max(np.abs(hh - h)) < 1e-11 * np.linalg.norm(vv)
    assert abs(nrm.value - np.linalg.norm(w)) < 1e-11 * np.linalg.norm(vv)
    assert relerr(out, w) < 1e-12


def check_mix(dev, q=4, n=20011):
    T = so.table("T_inv", q)
    src = synth(q * n, 4).reshape(q, n)
    dst0 = synth(q * n, 5).reshape(q, n)
    pT, _ = capi.darr(T)
    with capi.Context(dev) as ctx:
        ds, dd = ctx.upload(src), ctx.upload(dst0)
        ctx.call("spirk_mix", q, q, dd, n, ds, n, n, pT, 0, 1e-12)
        assert relerr(ctx.download(dd, (q, n)), so.mix(T, src)) < 1e-13
        ctx.call("spirk_vec_copy", dd, ctx.upload(dst0), q * n)
        ctx.call("spirk_mix", q, q, dd, n, ds, n, n, pT, 1, 1e-12)
        assert relerr(ctx.download(dd, (q, n)), dst0 + so.mix(T, src)) < 1e-13
        # rectangular: 1 x q (update u += tau * b^T k)
        b = so.table("b_vec_", q)
        pb, _ = capi.darr(0.1 * b)
        du = ctx.upload(dst0[0])
        ctx.call("spirk_mix", 1, q, du, n, ds, n, n, pb, 1, 0.0)
        assert relerr(ctx.download(du, (n,)), dst0[0] + 0.1 * (b @ src)) < 1e-13


def check_mix_peer_virtual(dev, R, m, n=30011):
    """spirk_mix_peer (gather formulation) and spirk_mix_peer_a2a_contract / _finish (all-to-all formulation) with the R
    ranks of a stage group emulated on one device (spirk_xbuf_create_virtual_group): every rank's result must equal its
    rows of the plain mixing so.mix (reference perform_basis_change, main.cc:1486-1534, incl. the 1e-12 cut-off)."""
    q = R * m
    rng = np.random.default_rng(q)
    T = so.table("T_inv", q) if 2 <= q <= 10 else rng.standard_normal((q, q))
    T = np.ascontiguousarray(T, dtype=np.float64)
    src = synth(q * n, 7).reshape(q, n)      # block j lives on rank j // m
    dst0 = synth(q * n, 8).reshape(q, n)
    ref = so.mix(T, src)
    with capi.Context(dev) as ctx:
        xs = (C.c_void_p * R)()
        ctx.call("spirk_xbuf_create_virtual_group", R, m * n, xs)
        try:
            for r in range(R):
                loc = C.c_void_p(dev.lib.spirk_comm_xbuf_local(xs[r]))
                ctx.call("spirk_copy_h2d", loc, np.ascontiguousarray(src[r * m:(r + 1) * m]).ctypes.data_as(C.c_void_p), m * n)
            for add in (0, 1):
                # gather formulation: rank r contracts all q blocks for its m rows
                for r in range(R):
                    d = ctx.upload(dst0[r * m:(r + 1) * m])
                    rows, _k = capi.darr(T[r * m:(r + 1) * m])
                    ctx.call("spirk_mix_peer", None, xs[r], m, m, d, n, n, rows, add, 1e-12)
                    want = ref[r * m:(r + 1) * m] + (dst0[r * m:(r + 1) * m] if add else 0.0)
                    assert relerr(ctx.download(d, (m, n)), want) < 1e-13, (R, m, r, add, "gather")
                    ctx.free(d)
                # all-to-all formulation: every rank contracts its chunk for all outputs, then every rank collects
                full, _f = capi.darr(T)
                for r in range(R):
                    ctx.call("spirk_mix_peer_a2a_contract", xs[r], m, n, full, 1e-12)
                for r in range(R):
                    d = ctx.upload(dst0[r * m:(r + 1) * m])
                    ctx.call("spirk_mix_peer_a2a_finish", xs[r], m, d, n, n, add)
                    want = ref[r * m:(r + 1) * m] + (dst0[r * m:(r + 1) * m] if add else 0.0)
                    assert relerr(ctx.download(d, (m, n)), want) < 1e-13, (R, m, r, add, "a2a")
                    ctx.free(d)
        finally:
            for r in range(R):
                ctx.call("spirk_comm_xbuf_destroy", xs[r])


def check_problem(dev, dim, k, r):
    lvl, olv = make_level(dim, k, r)
    prob = so.Problem(dim, k, r)
    with capi.Context(dev) as ctx:
        d = ctx.alloc(olv.N)
        ctx.call("spirk_problem_rhs_spatial", C.byref(lvl), d)
        assert relerr(ctx.download(d, (1,) + olv.shape), prob.rspace) < 1e-12
        ctx.call("spirk_problem_interpolate_solution", C.byref(lvl), d, 0.3)
        u = ctx.download(d, (1,) + olv.shape)
        assert relerr(u, prob.exact_nodal(0.3)) < 1e-12
        l2, li = C.c_double(), C.c_double()
        ctx.call("spirk_problem_error_norms", C.byref(lvl), d, 0.3, C.byref(l2), C.byref(li))
        e = prob.errors(u, 0.3)
        assert abs(l2.value - e[0]) < 1e-9 * e[0] and abs(li.value - e[1]) < 1e-9 * e[1]
        ctx.call("spirk_constraints_set_zero", C.byref(lvl), 1, d, olv.N)
        u2 = ctx.download(d, (1,) + olv.shape)
        assert np.all(u2[:, olv.bmask] == 0.0) and np.array_equal(u2[:, ~olv.bmask], u[:, ~olv.bmask])


def check_dense_matvec(dev, n=27, nb=3):
    rng = np.random.default_rng(0)
    A = rng.standard_normal((n, n))
    x = rng.standard_normal((nb, 40))
    with capi.Context(dev) as ctx:
        dA, dx, dy = ctx.upload(A), ctx.upload(x), ctx.alloc(nb * 40)
        ctx.call("spirk_dense_matvec", n, nb, dy, dx, 40, dA, 0)
        y = ctx.download(dy, (nb, 40))
        assert relerr(y[:, :n], x[:, :n] @ A.T) < 1e-13
        As = rng.standard_normal((nb, n, n))
        dAs = ctx.upload(As)
        ctx.call("spirk_dense_matvec", n, nb, dy, dx, 40, dAs, n * n)
        y = ctx.download(dy, (nb, 40))
        assert relerr(y[:, :n], np.einsum("bij,bj->bi", As, x[:, :n])) < 1e-13


D4 = [16.0, 3.1618475338398158, 2.9418686642961562, 5.644106850167844]
OP_CASES = [
    ("real", [16.0], [0.1]),
    ("real", [1.0], [0.0]),
    ("real", [0.0], [-1.0]),
    ("real", D4, [0.1]),
    ("coupled", [[5.0, 3.0], [-3.0, 5.0]], [0.1]),                       # complex pair
    ("coupled", so.table("A_inv", 4).tolist(), [0.1]),                    # IRK system matrix
]


def run_all(dev, small=True):
    for case in OP_CASES:
        check_op_apply(dev, 3, 4, 1 if small else 3, case)
    check_op_apply(dev, 2, 2, 3, OP_CASES[0])
    check_op_apply(dev, 2, 4, 2, OP_CASES[3])
    check_op_apply(dev, 3, 1, 2, OP_CASES[0])
    check_op_apply(dev, 3, 2, 2, OP_CASES[4])
    check_op_apply(dev, 3, 3, 1, OP_CASES[0])
    check_residual_and_cheb(dev, 3, 4, 1)
    check_residual_and_cheb(dev, 2, 2, 3, nb=1)
    check_coupled_fused(dev, 3, 4, 1)
    check_coupled_fused(dev, 2, 2, 3, ("coupled", [[2.0, 0.5], [-0.25, 3.0]], [0.1, 0.2]))
    check_apply_km(dev, 3, 4, 1)
    check_apply_km(dev, 2, 2, 3, nb=1)
    for (dim, k, r) in [(3, 4, 1), (2, 2, 3), (3, 1, 2)]:
        check_inverse_diagonal(dev, dim, k, r)
        check_transfer(dev, dim, k, r)
        check_problem(dev, dim, k, r)
    check_assemble_dense(dev, 3, 4)
    check_assemble_dense(dev, 2, 2)
    check_vector_ops(dev)
    check_mgs(dev)
    check_mix(dev)
    check_dense_matvec(dev)
