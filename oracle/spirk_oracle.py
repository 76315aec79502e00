"""
ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

Independent NumPy/SciPy restatement of the hot path of peterrum/dealii-spirk (stage-parallel
fully implicit Runge-Kutta for the heat equation).  It is used only by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline leg, as the checker.

PARITY UNPINNED: the reference ships no tests, golden vectors or logs, and its
arithmetic lives in deal.II (>= 9.3, unpinned, not vendored, not installable here) + Trilinos ML.
This file restates the published deal.II algorithms the reference composes
and is anchored on the reference's own call sites, cited per function as
`ref <file>:<line>` (paths relative to the reference root).  The independent anchor is the
sparse-direct solve of the stage system (`direct_irk_step`) and the manufactured solution.

Design: on the uniformly refined hypercube with a lexicographic DoF numbering every operator
of the reference is a Kronecker expression of assembled 1-D matrices.  The oracle applies the
*assembled* 1-D matrices (a different evaluation order from the product's per-cell CUDA
kernels and from the C cell-loop port in oracle/cpu_abi.cc — that is the point of an oracle).

Vectors are numpy arrays of shape (nb, n1, ..., n1) (nb = number of blocks: 1 for a scalar
vector, q for a stage block vector, 2 for a complex pair); the last axis is x.
"""
import os

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp
import scipy.sparse.linalg as spla

TABLE_FILE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                          "dealii_spirk_b200", "tables", "butcher_tables.txt")


# ----------------------------------------------------------------------------------------
# Butcher tables (ref main.cc:599-656 loaders; consumers 676-681, 1778-1786)
# ----------------------------------------------------------------------------------------
_tables_cache = None


def load_tables():
    global _tables_cache
    if _tables_cache is None:
        t = {}
        for line in open(TABLE_FILE):
            if line.startswith("#") or not line.strip():
                continue
            tok = line.split()
            label, q, m, n = tok[0], int(tok[1]), int(tok[2]), int(tok[3])
            v = np.array([float(x) for x in tok[4:]])
            t[(label, q)] = v.reshape(m, n) if m > 1 else v.copy()
        _tables_cache = t
    return _tables_cache


def table(label, q):
    return load_tables()[(label, q)]


# ----------------------------------------------------------------------------------------
# 1-D finite element building blocks: FE_Q(k) on Gauss-Lobatto points, QGauss(k+1)
# (ref main.cc:3028-3029)
# ----------------------------------------------------------------------------------------
def gll_nodes(k):
    if k == 1:
        return np.array([0.0, 1.0])
    L = np.polynomial.legendre
    c = np.zeros(k + 1)
    c[k] = 1.0
    dc = L.legder(c)
    x = np.sort(np.real(L.legroots(dc)))
    for _ in range(4):
        x = x - L.legval(x, dc) / L.legval(x, L.legder(dc))
    x = np.concatenate([[-1.0], x, [1.0]])
    x = 0.5 * (x - x[::-1])
    return 0.5 * (x + 1.0)


def gauss_rule(n):
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def lagrange_eval(nodes, x):
    """values B[q,i] = l_i(x_q) and derivatives D[q,i] = l_i'(x_q)."""
    n = len(nodes)
    x = np.atleast_1d(np.asarray(x, dtype=float))
    B = np.ones((len(x), n))
    D = np.zeros((len(x), n))
    for i in range(n):
        for j in range(n):
            if j != i:
                B[:, i] *= (x - nodes[j]) / (nodes[i] - nodes[j])
        for m in range(n):
            if m == i:
                continue
            t = np.ones(len(x)) / (nodes[i] - nodes[m])
            for j in range(n):
                if j != i and j != m:
                    t *= (x - nodes[j]) / (nodes[i] - nodes[j])
            D[:, i] += t
    return B, D


def element_matrices(k):
    """reference-cell 1-D mass and stiffness matrices with QGauss(k+1)."""
    nodes = gll_nodes(k)
    xq, wq = gauss_rule(k + 1)
    B, D = lagrange_eval(nodes, xq)
    Mh = B.T @ (wq[:, None] * B)
    Kh = D.T @ (wq[:, None] * D)
    return nodes, Mh, Kh


def prolongation_1d(k):
    """(2k+1) x (k+1): coarse GLL Lagrange polynomials at the fine nodes of the two children
    (ref preconditioner.h:267-271 sets the transfer up)."""
    nodes = gll_nodes(k)
    fine = np.concatenate([0.5 * nodes, 0.5 + 0.5 * nodes[1:]])
    B, _ = lagrange_eval(nodes, fine)
    B[np.abs(B) < 1e-15] = 0.0
    return B


class Level:
    """One uniformly refined hypercube level: 2^r cells per direction, FE_Q(k)."""

    def __init__(self, dim, k, r):
        self.dim, self.k, self.r = dim, k, r
        self.nc = 2 ** r
        self.n1 = k * self.nc + 1
        self.h = 1.0 / self.nc
        self.nodes, Mh, Kh = element_matrices(k)
        n1 = self.n1
        M1 = np.zeros((n1, n1))
        K1 = np.zeros((n1, n1))
        for c in range(self.nc):
            s = slice(c * k, c * k + k + 1)
            M1[s, s] += self.h * Mh
            K1[s, s] += Kh / self.h
        self.M1, self.K1 = M1, K1
        self.shape = (n1,) * dim
        self.N = n1 ** dim
        b1 = np.zeros(n1, dtype=bool)
        b1[0] = b1[-1] = True
        if dim == 2:
            self.bmask = b1[:, None] | b1[None, :]
        else:
            self.bmask = b1[:, None, None] | b1[None, :, None] | b1[None, None, :]
        self.x1 = np.concatenate([[0.0]] + [(c + self.nodes[1:]) * self.h for c in range(self.nc)])

    def zeros(self, nb=1):
        return np.zeros((nb,) + self.shape)

    def _ax(self, A, u, axis):
        # apply the 1-D matrix A along spatial `axis` (0 = x = last array axis)
        a = u.ndim - 1 - axis
        return np.moveaxis(np.tensordot(A, u, axes=(1, a)), 0, a)

    def apply(self, u, mass, lap):
        """(mass M + lap K) on unconstrained DoFs, identity on Dirichlet DoFs
        (ref operator.h:298-310, 379-421).  mass / lap may be arrays (nb,)."""
        v = u.copy()
        v[:, self.bmask] = 0.0
        # (arithmetic in the precision of u: float64 everywhere except the FP32 V-cycle study, GMG.fp32)
        mass = np.asarray(mass, dtype=u.dtype).reshape((-1,) + (1,) * self.dim)
        lap = np.asarray(lap, dtype=u.dtype).reshape((-1,) + (1,) * self.dim)
        M, K = self.M1.astype(u.dtype, copy=False), self.K1.astype(u.dtype, copy=False)
        a = self._ax(M, v, 0)
        b = self._ax(K, v, 0)
        if self.dim == 2:
            out = self._ax(M, mass * a + lap * b, 1) + lap * self._ax(K, a, 1)
        else:
            c = self._ax(M, a, 1)
            d = self._ax(K, a, 1)
            e = self._ax(M, b, 1)
            out = self._ax(M, mass * c + lap * (d + e), 2) + lap * self._ax(K, c, 2)
        out[:, self.bmask] = u[:, self.bmask]
        return out

    def diagonal(self, mass, lap):
        """assembled diagonal of mass M + lap K (ref operator.h:361-373 before inversion)."""
        m, kd = np.diag(self.M1), np.diag(self.K1)
        if self.dim == 2:
            return mass * np.multiply.outer(m, m) + lap * (np.multiply.outer(kd, m) + np.multiply.outer(m, kd))
        mm = np.multiply.outer
        return mass * mm(mm(m, m), m) + lap * (mm(mm(kd, m), m) + mm(mm(m, kd), m) + mm(mm(m, m), kd))

    def inverse_diagonal(self, mass, lap):
        """ref operator.h:361-373: abs(d) > 1e-10 ? 1/d : 1.0; constrained entries -> 1.0 (A4)."""
        d = self.diagonal(mass, lap).copy()
        d[self.bmask] = 0.0
        out = np.ones_like(d)
        nz = np.abs(d) > 1.0e-10
        out[nz] = 1.0 / d[nz]
        return out[None]


# ----------------------------------------------------------------------------------------
# operators (ref include/operator.h).  Each exposes vmult(u) and inverse_diagonal().
# ----------------------------------------------------------------------------------------
class ScalarOp:
    """MassLaplaceOperatorMatrixFree, ref operator.h:250-460 (coefficients are mutable)."""

    def __init__(self, level, mass=1.0, lap=1.0):
        self.level, self.mass, self.lap, self.nb = level, mass, lap, 1

    def reinit(self, mass, lap):
        self.mass, self.lap = mass, lap

    def vmult(self, u, mass=None, lap=None):
        if mass is not None:
            self.reinit(mass, lap)
        return self.level.apply(u, self.mass, self.lap)

    def inverse_diagonal(self):
        return self.level.inverse_diagonal(self.mass, self.lap)


class BatchedOp:
    """BatchedMassLaplaceOperatorMatrixFree, ref operator.h:749-881: block b gets d_b M + tau K."""

    def __init__(self, level, d_vec, tau=1.0):
        self.level, self.d_vec, self.tau, self.nb = level, np.asarray(d_vec, float), tau, len(d_vec)

    def vmult(self, u):
        return self.level.apply(u, self.d_vec, np.full(self.nb, self.tau))

    def inverse_diagonal(self):
        return np.concatenate([self.level.inverse_diagonal(d, self.tau) for d in self.d_vec])


class ComplexOp:
    """ComplexMassLaplaceOperatorMatrixFree, ref operator.h:529-698:
    [[l_re M + tau K, -l_im M], [l_im M, l_re M + tau K]], identity on constrained DoFs."""

    def __init__(self, level, lre=1.0, lim=1.0, tau=1.0):  # defaults ref operator.h:468-472
        self.level, self.lre, self.lim, self.tau, self.nb = level, lre, lim, tau, 2

    def reinit(self, lre, lim, tau):
        self.lre, self.lim, self.tau = lre, lim, tau

    def vmult(self, u):
        lv = self.level
        A = lv.apply(u, self.lre, self.tau)  # diagonal blocks incl. identity on boundary
        v = u.copy()
        v[:, lv.bmask] = 0.0
        Mv = lv.apply(v, 1.0, 0.0)
        Mv[:, lv.bmask] = 0.0
        out = A
        out[0] -= self.lim * Mv[1]
        out[1] += self.lim * Mv[0]
        return out

    def inverse_diagonal(self):
        # ref operator.h:560-575: real-part diagonal copied to both blocks
        d = self.level.inverse_diagonal(self.lre, self.tau)
        return np.concatenate([d, d])


# ----------------------------------------------------------------------------------------
# deal.II Krylov solvers
# ----------------------------------------------------------------------------------------
class NoConvergence(Exception):
    pass


class SolverControl:
    """SolverControl(n, tol) / ReductionControl(n, tol, reduce)."""

    def __init__(self, max_steps, tol, reduce=None):
        self.max_steps, self.tol, self.reduce = max_steps, tol, reduce
        self.last_step, self.last_value, self.initial = 0, 0.0, 0.0

    def check(self, step, value):
        self.last_step, self.last_value = step, value
        if step == 0:
            self.initial = value
        if self.reduce is not None and value <= self.reduce * self.initial:
            return "success"
        if value <= self.tol:
            return "success"
        if step >= self.max_steps or np.isnan(value):
            return "failure"
        return "iterate"


def dot(a, b):
    return float(np.vdot(a, b))


def solver_cg(A, x, b, P, control, lanczos=None):
    """deal.II SolverCG::solve.  A, P: callables.  Returns x.
    lanczos: optional dict collecting the Lanczos tridiagonal (diag, offdiag)."""
    if np.any(x != 0.0):
        g = A(x) - b
    else:
        g = -b
    res = np.sqrt(dot(g, g))
    conv = control.check(0, res)
    if conv != "iterate":
        if conv == "failure":
            raise NoConvergence()
        return x
    h = P(g)
    d = -h
    gh = dot(g, h)
    it = 0
    diag, off = [], []
    eigen_beta_alpha = 0.0
    old_alpha = beta = 0.0
    while True:
        it += 1
        h = A(d)
        alpha = gh / dot(d, h)
        x = x + alpha * d
        g = g + alpha * h
        res = np.sqrt(abs(dot(g, g)))
        if it > 1 and lanczos is not None:
            diag.append(1.0 / old_alpha + eigen_beta_alpha)
            eigen_beta_alpha = beta / old_alpha
            off.append(np.sqrt(beta) / old_alpha)
            lanczos["diag"], lanczos["off"] = list(diag), list(off)
        conv = control.check(it, res)
        if conv != "iterate":
            break
        h = P(g)
        beta = gh
        gh = dot(g, h)
        beta = gh / beta
        d = beta * d - h
        old_alpha = alpha
    if conv == "failure":
        raise NoConvergence(x)
    return x


def solver_gmres(A, x, b, P, control, max_n_tmp_vectors=30):
    """deal.II SolverGMRES::solve, default AdditionalData: left preconditioning, restart
    length max_n_tmp_vectors-2, stopping on the preconditioned residual, modified Gram-Schmidt
    with Kelley re-orthogonalisation test every 5th iteration."""
    n_tmp = max_n_tmp_vectors
    acc = 0
    re_orth = False
    state = "iterate"
    sqrt_eps = np.sqrt(np.finfo(float).eps)
    while state == "iterate":
        p = b - A(x)
        v = P(p)
        rho = np.sqrt(dot(v, v))
        state = control.check(acc, rho)
        if state != "iterate":
            break
        gamma = np.zeros(n_tmp)
        ci = np.zeros(n_tmp - 1)
        si = np.zeros(n_tmp - 1)
        H = np.zeros((n_tmp, n_tmp - 1))
        gamma[0] = rho
        V = [v / rho]
        dim = 0
        inner = 0
        while inner < n_tmp - 2 and state == "iterate":
            acc += 1
            vv = P(A(V[inner]))
            dim = inner + 1
            h = np.zeros(n_tmp)
            consider = (not re_orth) and (acc % 5 == 0)
            if consider:
                norm_start = np.sqrt(dot(vv, vv))
            h[0] = dot(vv, V[0])
            for i in range(1, dim):
                vv = vv - h[i - 1] * V[i - 1]
                h[i] = dot(vv, V[i])
            vv = vv - h[dim - 1] * V[dim - 1]
            norm_vv = np.sqrt(dot(vv, vv))
            if consider and not (norm_vv > 10.0 * norm_start * sqrt_eps):
                re_orth = True
            if re_orth:
                htmp = dot(vv, V[0])
                h[0] += htmp
                for i in range(1, dim):
                    vv = vv - htmp * V[i - 1]
                    htmp = dot(vv, V[i])
                    h[i] += htmp
                vv = vv - htmp * V[dim - 1]
                norm_vv = np.sqrt(dot(vv, vv))
            s = norm_vv
            h[inner + 1] = s
            if s != 0:
                vv = vv / s
            V.append(vv)
            # Givens rotation
            col = inner
            for i in range(col):
                dummy = h[i]
                h[i] = ci[i] * dummy + si[i] * h[i + 1]
                h[i + 1] = -si[i] * dummy + ci[i] * h[i + 1]
            rr = 1.0 / np.sqrt(h[col] * h[col] + h[col + 1] * h[col + 1])
            si[col] = h[col + 1] * rr
            ci[col] = h[col] * rr
            h[col] = ci[col] * h[col] + si[col] * h[col + 1]
            gamma[col + 1] = -si[col] * gamma[col]
            gamma[col] *= ci[col]
            H[:dim, inner] = h[:dim]
            rho = abs(gamma[dim])
            state = control.check(acc, rho)
            inner += 1
        y = sla.solve_triangular(H[:dim, :dim], gamma[:dim])
        for i in range(dim):
            x = x + y[i] * V[i]
    if state == "failure":
        raise NoConvergence(x)
    return x


# ----------------------------------------------------------------------------------------
# Chebyshev smoother (deal.II PreconditionChebyshev; configured
# ref preconditioner.h:219-223, 353-373)
# ----------------------------------------------------------------------------------------
class Chebyshev:
    def __init__(self, op, inv_diag, degree=5, smoothing_range=20.0, eig_cg_n_iterations=20):
        self.op, self.dinv, self.degree = op, inv_diag, degree
        nb = inv_diag.shape[0]
        shape = inv_diag.shape[1:]
        N = int(np.prod(shape))
        # set_initial_guess: (global index % 11) minus mean, per block
        g = (np.arange(N) % 11).astype(float).reshape(shape)
        g = g - g.mean()
        v0 = np.repeat(g[None], nb, axis=0)
        control = SolverControl(eig_cg_n_iterations, np.sqrt(np.finfo(float).eps), 1e-10)
        lz = {}
        try:
            solver_cg(op.vmult, np.zeros_like(v0), v0, lambda r: self.dinv * r, control, lanczos=lz)
        except NoConvergence:
            pass
        if lz.get("diag"):
            n = len(lz["diag"])
            ev = sla.eigvalsh_tridiagonal(np.array(lz["diag"]), np.array(lz["off"][: n - 1])) if n > 1 \
                else np.array(lz["diag"])
            self.min_ev, self.max_ev = float(ev[0]), 1.2 * float(ev[-1])
        else:
            self.min_ev = self.max_ev = 1.0
        self.cg_iterations = control.last_step
        alpha = self.max_ev / smoothing_range if smoothing_range > 1.0 else min(0.9 * self.max_ev, self.min_ev)
        self.delta = 0.5 * (self.max_ev - alpha)
        self.theta = 0.5 * (self.max_ev + alpha)

    def vmult(self, b):
        dinv = self.dinv.astype(b.dtype, copy=False)
        x = (1.0 / self.theta) * (dinv * b)
        if self.degree < 2 or abs(self.delta) < 1e-40:
            return x
        xold = None
        rhok, sigma = self.delta / self.theta, self.theta / self.delta
        for k in range(self.degree - 1):
            Ax = self.op.vmult(x)
            rhokp = 1.0 / (2.0 * sigma - rhok)
            f1, f2 = rhokp * rhok, 2.0 * rhokp / self.delta
            rhok = rhokp
            if k == 0:
                xn = (1.0 + f1) * x + f2 * dinv * (b - Ax)
            else:
                xn = (1.0 + f1) * x - f1 * xold + f2 * dinv * (b - Ax)
            xold, x = x, xn
        return x


# ----------------------------------------------------------------------------------------
# geometric multigrid (ref preconditioner.h:219-501)
# ----------------------------------------------------------------------------------------
class GMG:
    """PreconditionerGMG: one V-cycle, Chebyshev(5) pre/post smoothing on every level, levels
    0..r (level 0 = one cell), exact coarse solve for scalar vectors (Trilinos ML on a
    (k+1)^d matrix degenerates to a direct solve, A10), level-0 smoother for block vectors."""

    def __init__(self, dim, k, r, make_op, block=False):
        self.levels = [Level(dim, k, l) for l in range(r + 1)]
        self.ops = [make_op(lv) for lv in self.levels]
        self.block = block
        self.k = k
        P = prolongation_1d(k)
        self.P1 = []
        for l in range(1, r + 1):
            nc_c = 2 ** (l - 1)
            Pg = np.zeros((self.levels[l].n1, self.levels[l - 1].n1))
            for c in range(nc_c):
                Pg[2 * k * c: 2 * k * c + 2 * k + 1, k * c: k * c + k + 1] = P
            self.P1.append(Pg)
        self.smoothers = None
        # FP32 V-cycle study (reference preconditioner.h:120-142 anticipates a float level vector): the
        # set-up (diagonals, eigenvalue estimates, coarse inverse) stays FP64, one V-cycle runs in float32 and its result
        # is widened again for the FP64 outer Krylov solver
        self.fp32 = False

    def reinit(self, ops_for_setup=None):
        """ref preconditioner.h:341-447.  ops_for_setup lets the caller replicate SURVEY 2.4(9):
        diagonals / eigenvalue estimates computed with different coefficients than vmult."""
        setup_ops = ops_for_setup if ops_for_setup is not None else self.ops
        self.smoothers = []
        for lv, op, sop in zip(self.levels, self.ops, setup_ops):
            dinv = sop.inverse_diagonal()
            sm = Chebyshev(sop, dinv)
            sm.op = op  # the V-cycle applies the live level operator
            self.smoothers.append(sm)
        if not self.block:
            lv0, op0 = self.levels[0], setup_ops[0]
            ni = lv0.N - int(lv0.bmask.sum())
            if ni > 0:
                A = np.zeros((ni, ni))
                idx = np.argwhere(~lv0.bmask)
                for j, ij in enumerate(idx):
                    e = lv0.zeros()
                    e[(0,) + tuple(ij)] = 1.0
                    A[:, j] = op0.vmult(e)[0][~lv0.bmask]
                self.coarse_inv = np.linalg.inv(A)
            else:
                self.coarse_inv = None

    def prolongate(self, l, uc):
        lvc = self.levels[l - 1]
        v = uc.copy()
        v[:, lvc.bmask] = 0.0
        Pg = self.P1[l - 1].astype(uc.dtype, copy=False)
        out = v
        for ax in range(lvc.dim):
            a = out.ndim - 1 - ax
            out = np.moveaxis(np.tensordot(Pg, out, axes=(1, a)), 0, a)
        return out

    def restrict(self, l, uf):
        lvc = self.levels[l - 1]
        Pg = self.P1[l - 1].astype(uf.dtype, copy=False)
        out = uf
        for ax in range(lvc.dim):
            a = out.ndim - 1 - ax
            out = np.moveaxis(np.tensordot(Pg.T, out, axes=(1, a)), 0, a)
        out = out.copy()
        out[:, lvc.bmask] = 0.0
        return out

    def coarse(self, b):
        if self.block:
            return self.smoothers[0].vmult(b)
        lv0 = self.levels[0]
        x = b.copy()  # identity rows on constrained DoFs
        if self.coarse_inv is not None:
            x[0][~lv0.bmask] = self.coarse_inv.astype(b.dtype, copy=False) @ b[0][~lv0.bmask]
        return x

    def vcycle(self, l, defect):
        if l == 0:
            return self.coarse(defect)
        op, sm = self.ops[l], self.smoothers[l]
        x = sm.vmult(defect)                      # pre-smoothing, zero start (A8)
        t = defect - op.vmult(x)
        dc = self.restrict(l, t)
        xc = self.vcycle(l - 1, dc)
        x = x + self.prolongate(l, xc)
        r = defect - op.vmult(x)                  # post-smoothing: x += S (b - A x)
        x = x + sm.vmult(r)
        return x

    def vmult(self, src):
        if self.fp32:
            return self.vcycle(len(self.levels) - 1, src.astype(np.float32)).astype(np.float64)
        return self.vcycle(len(self.levels) - 1, src)


# ----------------------------------------------------------------------------------------
# the heat-equation problem (ref main.cc:3014-3603)
# ----------------------------------------------------------------------------------------
def time_factor_rhs(t, dim):
    """g(t) of the separable forcing f = s(x) g(t) (ref main.cc:3523-3539), a=2, a_t=.5, c_t=1."""
    pi = np.pi
    return (pi * np.cos(pi * t) - 0.5 * (np.sin(pi * t) + 1) + dim * 4.0 * pi * pi * (np.sin(pi * t) + 1)) \
        * np.exp(-0.5 * t)


def time_factor_sol(t):
    return (1 + np.sin(np.pi * t)) * np.exp(-0.5 * t)


class Problem:
    def __init__(self, dim, k, r):
        self.dim, self.k, self.r = dim, k, r
        self.level = lv = Level(dim, k, r)
        # spatial load vector r_i = int phi_i s(x), QGauss(k+1) (ref main.cc:3213-3219, A12)
        xq, wq = gauss_rule(k + 1)
        B, _ = lagrange_eval(lv.nodes, xq)
        r1 = np.zeros(lv.n1)
        for c in range(lv.nc):
            s = np.sin(2 * np.pi * (c + xq) * lv.h)
            r1[c * k: c * k + k + 1] += lv.h * (B.T @ (wq * s))
        self.r1 = r1
        rs = r1
        for _ in range(dim - 1):
            rs = np.multiply.outer(rs, r1)
        rs = rs.copy()
        rs[lv.bmask] = 0.0
        self.rspace = rs[None]
        # error quadrature QGauss(k+2) (ref main.cc:3436-3462)
        xe, we = gauss_rule(k + 2)
        Be, _ = lagrange_eval(lv.nodes, xe)
        E = np.zeros((lv.nc * (k + 2), lv.n1))
        for c in range(lv.nc):
            E[c * (k + 2):(c + 1) * (k + 2), c * k: c * k + k + 1] = Be
        self.E = E
        self.xe = np.concatenate([(c + xe) * lv.h for c in range(lv.nc)])
        self.we = np.concatenate([we * lv.h for _ in range(lv.nc)])

    def rhs(self, t):
        return time_factor_rhs(t, self.dim) * self.rspace

    def exact_nodal(self, t):
        s = np.sin(2 * np.pi * self.level.x1)
        u = s
        for _ in range(self.dim - 1):
            u = np.multiply.outer(u, s)
        return (time_factor_sol(t) * u)[None]

    def initial(self):
        u = self.exact_nodal(0.0)
        u[:, self.level.bmask] = 0.0
        return u

    def errors(self, u, t):
        """(L2, Linf) error vs the analytical solution with QGauss(k+2) (ref main.cc:3436-3469)."""
        uh = u[0]
        for ax in range(self.dim):
            a = uh.ndim - 1 - ax
            uh = np.moveaxis(np.tensordot(self.E, uh, axes=(1, a)), 0, a)
        s = np.sin(2 * np.pi * self.xe)
        ue, w = s, self.we
        for _ in range(self.dim - 1):
            ue = np.multiply.outer(ue, s)
            w = np.multiply.outer(w, self.we)
        d = uh - time_factor_sol(t) * ue
        return float(np.sqrt(np.sum(w * d * d))), float(np.max(np.abs(d)))


# ----------------------------------------------------------------------------------------
# time integrators (ref main.cc:450-2937)
# ----------------------------------------------------------------------------------------
def mix(T, blocks, cut=1e-12):
    """dst_i = sum_j T_ij src_j skipping |T_ij| <= cut (ref main.cc:1100-1104, 1164-1168,
    1511-1529).  blocks: (q, ...) array."""
    Tm = np.where(np.abs(T) > cut, T, 0.0)
    return np.tensordot(Tm, blocks, axes=(1, 0))


class IRK:
    """IRK / IRKStageParallel algebra (ref main.cc:771-1222 and 1229-1760; SURVEY B).
    batched=True reproduces `irk_batched` (one block GMG with a single Chebyshev range and a
    smoother coarse solve)."""

    def __init__(self, prob, q, tau, outer_tol=1e-8, inner_tol=0.0, batched=False):
        self.prob, self.q, self.tau = prob, q, tau
        self.outer_tol, self.inner_tol, self.batched = outer_tol, inner_tol, batched
        self.A_inv, self.T, self.T_inv = table("A_inv", q), table("T", q), table("T_inv", q)
        self.b, self.c, self.d = table("b_vec_", q), table("c_vec_", q), table("D_vec_", q)
        self.op = ScalarOp(prob.level)
        dim, k, r = prob.dim, prob.k, prob.r
        if batched:
            self.gmg = GMG(dim, k, r, lambda lv: BatchedOp(lv, self.d, tau), block=True)
            self.gmg.reinit()
        else:
            self.gmgs = []
            for i in range(q):
                g = GMG(dim, k, r, lambda lv, i=i: ScalarOp(lv, self.d[i], tau))
                g.reinit()
                self.gmgs.append(g)
        self.n_outer, self.n_inner = [], []

    def system_vmult(self, v):
        # ref main.cc:1014-1028 / 1580-1592: dst_i = tau K v_i + sum_j A_inv[i][j] M v_j
        lv = self.prob.level
        Kv = lv.apply(v, 0.0, self.tau)
        Mv = lv.apply(v, 1.0, 0.0)
        return Kv + np.tensordot(self.A_inv, Mv, axes=(1, 0))

    def precondition(self, src):
        # ref main.cc:1095-1173 / 1646-1707
        t = mix(self.T_inv, src)
        if self.batched:
            z = self.gmg.vmult(t)
            self._inner += 1
        else:
            z = np.zeros_like(t)
            for i in range(self.q):
                if self.inner_tol > 0.0:
                    ctl = SolverControl(100, 1e-10, self.inner_tol)
                    opi = self.gmgs[i].ops[-1]
                    z[i:i + 1] = solver_cg(opi.vmult, np.zeros_like(t[i:i + 1]), t[i:i + 1],
                                           self.gmgs[i].vmult, ctl)
                    self._inner += ctl.last_step
                else:
                    z[i:i + 1] = self.gmgs[i].vmult(t[i:i + 1])
                    self._inner += 1
        return mix(self.T, z)

    def rhs(self, u, time):
        # ref main.cc:867-891 / 1343-1349; `time` is t_{n+1}
        lv = self.prob.level
        tmp = lv.apply(u, 0.0, -1.0)
        g = np.concatenate([self.prob.rhs(time + (self.c[i] - 1.0) * self.tau) + tmp for i in range(self.q)])
        return np.tensordot(self.A_inv, g, axes=(1, 0))

    def step(self, u, time):
        rhs = self.rhs(u, time)
        self._inner = 0
        ctl = SolverControl(1000, 1e-20, self.outer_tol)
        ksol = solver_gmres(self.system_vmult, np.zeros_like(rhs), rhs, self.precondition, ctl)
        self.n_outer.append(ctl.last_step)
        self.n_inner.append(self._inner)
        self.stages = ksol
        unew = u + self.tau * np.tensordot(self.b, ksol, axes=(0, 0))[None]
        unew[:, self.prob.level.bmask] = 0.0   # constraints.distribute (ref main.cc:3355)
        return unew


class ComplexIRK:
    """ComplexIRK / ComplexSPIRK algebra (ref main.cc:1886-2375, 2382-2934; SURVEY B).
    batched=True: complex block GMG (`complex_*_batched`), else PRESB with real GMGs.
    literal_setup=True replicates SURVEY 2.4(9): complex GMG smoother set up with the
    constructor-default coefficients (1,1,1)."""

    def __init__(self, prob, q, tau, outer_tol=1e-8, inner_tol=0.0, batched=False, literal_setup=True):
        self.prob, self.q, self.tau = prob, q, tau
        self.outer_tol, self.inner_tol, self.batched = outer_tol, inner_tol, batched
        self.A_inv = table("A_inv", q)
        self.T_re, self.T_im = table("T_re", q), table("T_im", q)
        self.Ti_re, self.Ti_im = table("T_inv_re", q), table("T_inv_im", q)
        self.b, self.c = table("b_vec_", q), table("c_vec_", q)
        self.d_re, self.d_im = table("D_vec_re_", q), table("D_vec_im_", q)
        self.nred = (q + 1) // 2
        dim, k, r = prob.dim, prob.k, prob.r
        self.gmgs = []
        for i in range(self.nred):
            lre, lim = self.d_re[2 * i], self.d_im[2 * i]
            if batched:
                g = GMG(dim, k, r, lambda lv: ComplexOp(lv, lre, lim, tau), block=True)
                if literal_setup:
                    g.reinit([ComplexOp(lv, 1.0, 1.0, 1.0) for lv in g.levels])
                else:
                    g.reinit()
            else:
                g = GMG(dim, k, r, lambda lv: ScalarOp(lv, lre + lim, tau))
                g.reinit()
            self.gmgs.append(g)
        self.n_outer, self.n_inner = [], []

    def presb(self, i, src):
        # ref main.cc:2283-2335
        lv = self.prob.level
        lre, lim = self.d_re[2 * i], self.d_im[2 * i]
        g = self.gmgs[i]

        def H_inv(rhs):
            if self.inner_tol == 0.0:
                self._inner += 1
                return g.vmult(rhs)
            ctl = SolverControl(100, self.inner_tol)
            x = solver_cg(g.ops[-1].vmult, np.zeros_like(rhs), rhs, g.vmult, ctl)
            self._inner += ctl.last_step
            return x

        t0 = src[0:1] + src[1:2]
        x0 = H_inv(t0)
        t0 = src[1:2] - lv.apply(x0, lim, 0.0)
        x1 = H_inv(t0)
        return np.concatenate([x0 - x1, x1])

    def step(self, u, time):
        lv = self.prob.level
        q, tau = self.q, self.tau
        tmp = lv.apply(u, 0.0, -1.0)
        g = np.concatenate([self.prob.rhs(time + (self.c[i] - 1.0) * tau) + tmp for i in range(q)])
        rhs = np.tensordot(self.A_inv, g, axes=(1, 0))
        z = []
        outer, self._inner = [], 0
        for i in range(self.nred):
            src = np.stack([np.tensordot(self.Ti_re[2 * i], rhs, axes=(0, 0)),
                            np.tensordot(self.Ti_im[2 * i], rhs, axes=(0, 0))])
            lre, lim = self.d_re[2 * i], self.d_im[2 * i]
            opc = ComplexOp(lv, lre, lim, tau)
            ctl = SolverControl(1000, 1e-20, self.outer_tol)
            if self.batched:
                P = self.gmgs[i].vmult
            else:
                P = lambda s, i=i: self.presb(i, s)
            zi = solver_gmres(opc.vmult, np.zeros_like(src), src, P, ctl)
            outer.append(ctl.last_step)
            z.append(zi)
        self.n_outer.append(outer)
        self.n_inner.append(self._inner)
        ksol = np.zeros_like(rhs)
        for i in range(q):
            for j in range(self.nred):
                s = 2.0 if j < q // 2 else 1.0
                ksol[i] += s * self.T_re[i, 2 * j] * z[j][0] - s * self.T_im[i, 2 * j] * z[j][1]
        self.stages = ksol
        unew = u + tau * np.tensordot(self.b, ksol, axes=(0, 0))[None]
        unew[:, lv.bmask] = 0.0
        return unew


class OneStepTheta:
    """ref main.cc:476-595.  literal=True keeps the reference's signs (rhs (M+(1-th)tau K)u,
    matrix M - th tau K); literal=False is Crank-Nicolson with the correct signs."""

    def __init__(self, prob, tau, literal=False):
        self.prob, self.tau, self.theta, self.literal = prob, tau, 0.5, literal
        sgn = -1.0 if literal else 1.0
        self.mass, self.lap = 1.0, sgn * self.theta * tau
        self.gmg = GMG(prob.dim, prob.k, prob.r, lambda lv: ScalarOp(lv, self.mass, self.lap))
        self.gmg.reinit()
        self.n_iter = []

    def step(self, u, time):
        lv, th, tau = self.prob.level, self.theta, self.tau
        sgn = 1.0 if self.literal else -1.0
        rhs = lv.apply(u, 1.0, sgn * (1 - th) * tau)
        rhs = rhs + tau * th * self.prob.rhs(time) + tau * (1 - th) * self.prob.rhs(time - tau)
        ctl = SolverControl(1000, 1e-8 * np.sqrt(dot(rhs, rhs)))
        op = self.gmg.ops[-1]
        x = solver_cg(op.vmult, u.copy(), rhs, self.gmg.vmult, ctl)
        self.n_iter.append(ctl.last_step)
        x[:, lv.bmask] = 0.0
        return x


# ----------------------------------------------------------------------------------------
# independent anchor: sparse-direct solve of the stage system (no iterative code path shared)
# ----------------------------------------------------------------------------------------
def direct_irk_step(prob, q, tau, u, time):
    lv = prob.level
    n1, dim = lv.n1, lv.dim
    I = slice(1, n1 - 1)
    Mi, Ki = sp.csr_matrix(lv.M1[I, I]), sp.csr_matrix(lv.K1[I, I])
    if dim == 2:
        M = sp.kron(Mi, Mi)
        K = sp.kron(Ki, Mi) + sp.kron(Mi, Ki)
    else:
        M = sp.kron(sp.kron(Mi, Mi), Mi)
        K = sp.kron(sp.kron(Ki, Mi), Mi) + sp.kron(sp.kron(Mi, Ki), Mi) + sp.kron(sp.kron(Mi, Mi), Ki)
    A_inv, b, c = table("A_inv", q), table("b_vec_", q), table("c_vec_", q)
    inner = (slice(None),) + (I,) * dim
    ui = u[inner].reshape(-1)
    Ku = K @ ui
    g = np.stack([prob.rhs(time + (c[i] - 1.0) * tau)[inner].reshape(-1) - Ku for i in range(q)])
    rhs = (A_inv @ g).reshape(-1)
    key = ("lu", q, tau)
    cache = prob.__dict__.setdefault("_direct_cache", {})
    if key not in cache:  # the stage matrix does not change between steps
        S = (sp.kron(sp.csr_matrix(A_inv), M) + tau * sp.kron(sp.identity(q), K)).tocsc()
        cache[key] = spla.splu(S)
    ksol = cache[key].solve(rhs).reshape(q, -1)
    unew = u.copy()
    ni = n1 - 2
    unew[inner] = (ui + tau * (b @ ksol)).reshape((1,) + (ni,) * dim)
    stages = np.zeros((q,) + lv.shape)
    stages[inner] = ksol.reshape((q,) + (ni,) * dim)
    return unew, stages


def set_fp32_vcycle(integ, on=True):
    """switch every multigrid preconditioner of a time integrator to the float32 V-cycle (study only)"""
    n = 0
    for v in vars(integ).values():
        for g in (v if isinstance(v, (list, tuple)) else [v]):
            if isinstance(g, GMG):
                g.fp32 = on
                n += 1
    return n


def run(scheme, dim, k, r, q, tau, end_time, outer_tol=1e-8, inner_tol=0.0, fp32_vcycle=False, **kw):
    """mirror of Problem::run's time loop (ref main.cc:3298-3358).  Returns a dict of per-step
    records: errors, iteration counts, nodal l2 norm."""
    prob = Problem(dim, k, r)
    u = prob.initial()
    if scheme == "ost":
        integ = OneStepTheta(prob, tau, **kw)
    elif scheme in ("irk", "spirk", "irk_batched"):
        integ = IRK(prob, q, tau, outer_tol, inner_tol, batched=(scheme == "irk_batched"))
    elif scheme.startswith("complex"):
        integ = ComplexIRK(prob, q, tau, outer_tol, inner_tol, batched=scheme.endswith("batched"), **kw)
    else:
        raise ValueError(scheme)
    if fp32_vcycle:
        assert set_fp32_vcycle(integ) > 0, "no multigrid preconditioner found"
    out = {"errors": [prob.errors(u, 0.0)], "norms": [], "prob": prob, "integ": integ}
    t = 0.0
    while end_time - t > 1e-4 * tau:
        t = t + tau
        u = integ.step(u, t)
        out["errors"].append(prob.errors(u, t))
        out["norms"].append(float(np.sqrt(dot(u, u))))
    out["u"] = u
    return out
