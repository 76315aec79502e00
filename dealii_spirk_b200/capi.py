"""ctypes bindings of include/spirk_b200.h (device layer) — thin, no logic.

`DeviceLib(path)` binds any shared library that implements the header.  The product library is
dealii_spirk_b200/libspirk_b200.so (CUDA, sm_100a); tests may also bind the CPU double in
oracle/_build/libspirk_cpu.so, but nothing in this package ever does.
"""
import ctypes as C
import os

import numpy as np

MAX_BLOCKS = 16
OP_REAL, OP_COUPLED = 0, 1
dp = C.POINTER(C.c_double)


class Level(C.Structure):
    _fields_ = [("dim", C.c_int), ("degree", C.c_int), ("n_cells_1d", C.c_int), ("slab", C.c_int)]

    @property
    def n1(self):
        return self.degree * self.n_cells_1d + 1

    @property
    def n_dofs(self):
        return self.n1 ** self.dim


class OpDesc(C.Structure):
    _fields_ = [("kind", C.c_int), ("nb", C.c_int), ("mass", C.c_double * MAX_BLOCKS),
                ("laplace", C.c_double * MAX_BLOCKS), ("coupling", C.c_double * (MAX_BLOCKS * MAX_BLOCKS))]


def real_op(mass, laplace):
    mass = np.atleast_1d(np.asarray(mass, float))
    laplace = np.broadcast_to(np.atleast_1d(np.asarray(laplace, float)), mass.shape)
    op = OpDesc()
    op.kind, op.nb = OP_REAL, len(mass)
    for i in range(op.nb):
        op.mass[i], op.laplace[i] = mass[i], laplace[i]
    return op


def coupled_op(coupling, laplace):
    coupling = np.asarray(coupling, float)
    nb = coupling.shape[0]
    laplace = np.broadcast_to(np.atleast_1d(np.asarray(laplace, float)), (nb,))
    op = OpDesc()
    op.kind, op.nb = OP_COUPLED, nb
    for i in range(nb):
        op.laplace[i] = laplace[i]
        for j in range(nb):
            op.coupling[i * nb + j] = coupling[i, j]
    return op


class SpirkError(RuntimeError):
    pass


_SIGS = {
    "spirk_ctx_create": [C.POINTER(C.c_void_p), C.c_int],
    "spirk_ctx_destroy": [C.c_void_p],
    "spirk_ctx_sync": [C.c_void_p],
    "spirk_ctx_timer_begin": [C.c_void_p],
    "spirk_ctx_timer_end": [C.c_void_p, dp],
    "spirk_ctx_set_option": [C.c_void_p, C.c_char_p, C.c_int],
    "spirk_graph_begin": [C.c_void_p],
    "spirk_graph_end": [C.c_void_p, C.POINTER(C.c_void_p)],
    "spirk_graph_launch": [C.c_void_p, C.c_void_p],
    "spirk_graph_destroy": [C.c_void_p],
    "spirk_malloc": [C.c_void_p, C.POINTER(C.c_void_p), C.c_size_t],
    "spirk_free": [C.c_void_p, C.c_void_p],
    "spirk_copy_h2d": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t],
    "spirk_copy_d2h": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t],
    "spirk_malloc_host": [C.c_void_p, C.POINTER(C.c_void_p), C.c_size_t],
    "spirk_free_host": [C.c_void_p, C.c_void_p],
    "spirk_op_apply": [C.c_void_p, C.POINTER(Level), C.POINTER(OpDesc), C.c_void_p, C.c_void_p, C.c_longlong],
    "spirk_op_apply_km": [C.c_void_p, C.POINTER(Level), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, dp, dp],
    "spirk_op_residual": [C.c_void_p, C.POINTER(Level), C.POINTER(OpDesc), C.c_void_p, C.c_void_p, C.c_void_p,
                          C.c_longlong],
    "spirk_op_cheb_step": [C.c_void_p, C.POINTER(Level), C.POINTER(OpDesc), C.c_void_p, C.c_void_p, C.c_void_p,
                           C.c_void_p, C.c_void_p, C.c_longlong, dp, dp],
    "spirk_op_cheb_first": [C.c_void_p, C.POINTER(Level), C.POINTER(OpDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong,
                            dp, dp, dp],
    "spirk_op_cheb_step_diag": [C.c_void_p, C.POINTER(Level), C.POINTER(OpDesc), C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, dp, dp, C.c_longlong, dp, dp],
    "spirk_op_cheb_first_diag": [C.c_void_p, C.POINTER(Level), C.POINTER(OpDesc), C.c_void_p, C.c_void_p, C.c_void_p, dp, dp,
                                 C.c_longlong, dp, dp, dp],
    "spirk_op_inverse_diagonal": [C.c_void_p, C.POINTER(Level), C.c_void_p, C.c_double, C.c_double],
    "spirk_op_assemble_dense": [C.c_void_p, C.POINTER(Level), C.c_double, C.c_double, C.c_void_p],
    "spirk_mg_prolongate_add": [C.c_void_p, C.POINTER(Level), C.c_int, C.c_void_p, C.c_longlong, C.c_void_p,
                                C.c_longlong],
    "spirk_mg_restrict": [C.c_void_p, C.POINTER(Level), C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong],
    "spirk_dense_matvec": [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong],
    "spirk_vec_set": [C.c_void_p, C.c_void_p, C.c_longlong, C.c_double],
    "spirk_vec_copy": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong],
    "spirk_vec_scale": [C.c_void_p, C.c_void_p, C.c_longlong, C.c_double],
    "spirk_vec_axpy": [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_longlong],
    "spirk_vec_sadd": [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_longlong],
    "spirk_vec_add2": [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_double, C.c_void_p, C.c_longlong],
    "spirk_vec_equ": [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_longlong],
    "spirk_vec_scale_pointwise": [C.c_void_p, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_longlong, dp],
    "spirk_vec_dot": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, dp],
    "spirk_vec_add_and_dot": [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_longlong, dp],
    "spirk_vec_sum": [C.c_void_p, C.c_void_p, C.c_longlong, dp],
    "spirk_vec_dot_strided": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_longlong, dp],
    "spirk_vec_sum_strided": [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_longlong, dp],
    "spirk_vec_add_and_dot_strided": [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int,
                                      C.c_longlong, dp],
    "spirk_gmres_mgs_strided": [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_longlong, C.c_int, C.c_longlong, dp, dp],
    "spirk_problem_error_norms_partial": [C.c_void_p, C.POINTER(Level), C.c_void_p, C.c_double, dp, dp],
    "spirk_comm_split": [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)],
    "spirk_comm_allreduce_max": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong],
    "spirk_halo_exchange": [C.c_void_p, C.c_void_p, C.POINTER(Level), C.c_int, C.c_void_p, C.c_longlong, C.c_int, C.c_int],
    "spirk_gmres_mgs": [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_longlong, dp, dp],
    "spirk_mix": [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_longlong,
                  dp, C.c_int, C.c_double],
    "spirk_problem_rhs_spatial": [C.c_void_p, C.POINTER(Level), C.c_void_p],
    "spirk_problem_interpolate_solution": [C.c_void_p, C.POINTER(Level), C.c_void_p, C.c_double],
    "spirk_problem_error_norms": [C.c_void_p, C.POINTER(Level), C.c_void_p, C.c_double, dp, dp],
    "spirk_constraints_set_zero": [C.c_void_p, C.POINTER(Level), C.c_int, C.c_void_p, C.c_longlong],
    "spirk_comm_unique_id": [C.c_char_p],
    "spirk_comm_create": [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)],
    "spirk_comm_destroy": [C.c_void_p],
    "spirk_comm_rank": [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)],
    "spirk_comm_allreduce_sum": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong],
    "spirk_comm_allgather": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong],
    "spirk_ctx_set_reduction_comm": [C.c_void_p, C.c_void_p],
    "spirk_comm_xbuf_create": [C.c_void_p, C.c_void_p, C.c_longlong, C.POINTER(C.c_void_p)],
    "spirk_comm_xbuf_destroy": [C.c_void_p, C.c_void_p],
    "spirk_mix_peer": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_longlong, dp,
                       C.c_int, C.c_double],
    "spirk_mix_peer_a2a": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_longlong, C.c_longlong, dp,
                           C.c_int, C.c_double],
    "spirk_mix_peer_a2a_contract": [C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, dp, C.c_double],
    "spirk_mix_peer_a2a_finish": [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_longlong, C.c_longlong, C.c_int],
    "spirk_xbuf_create_virtual_group": [C.c_void_p, C.c_int, C.c_longlong, C.POINTER(C.c_void_p)],
}
# every symbol include/spirk_b200.h declares (tests check the library exports all of them)
ALL_SYMBOLS = sorted(list(_SIGS) + ["spirk_backend", "spirk_last_error", "spirk_ctx_launch_count",
                                    "spirk_level_n_dofs", "spirk_comm_xbuf_local", "spirk_op_fuses_own_diagonal"])


class DeviceLib:
    def __init__(self, path):
        if not os.path.exists(path):
            raise SpirkError(f"{path} not found — run `python -c 'import __graft_entry__ as g; g.build()'`")
        self.path = path
        self.lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
        for name, sig in _SIGS.items():
            f = getattr(self.lib, name)
            f.argtypes, f.restype = sig, C.c_int
        self.lib.spirk_backend.restype = C.c_char_p
        self.lib.spirk_last_error.restype = C.c_char_p
        self.lib.spirk_ctx_launch_count.argtypes = [C.c_void_p]
        self.lib.spirk_ctx_launch_count.restype = C.c_longlong
        self.lib.spirk_level_n_dofs.argtypes = [C.POINTER(Level)]
        self.lib.spirk_level_n_dofs.restype = C.c_longlong
        self.lib.spirk_comm_xbuf_local.argtypes = [C.c_void_p]
        self.lib.spirk_comm_xbuf_local.restype = C.c_void_p

    def backend(self):
        return self.lib.spirk_backend().decode()

    def call(self, name, *args):
        st = getattr(self.lib, name)(*args)
        if st != 0:
            raise SpirkError(f"{name} failed with status {st}: {self.lib.spirk_last_error().decode()}")


class Context:
    """Small convenience layer for tests / bench: device buffers as opaque pointers + numpy staging."""

    def __init__(self, dev: DeviceLib, device=0):
        self.dev = dev
        h = C.c_void_p()
        dev.call("spirk_ctx_create", C.byref(h), device)
        self.h = h
        self._bufs = []

    def close(self):
        if self.h:
            for p in self._bufs:
                self.dev.call("spirk_free", self.h, p)
            self._bufs = []
            self.dev.call("spirk_ctx_destroy", self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def alloc(self, n):
        p = C.c_void_p()
        self.dev.call("spirk_malloc", self.h, C.byref(p), int(n))
        self._bufs.append(p)
        return p

    def free(self, p):
        self._bufs = [q for q in self._bufs if q.value != p.value]
        self.dev.call("spirk_free", self.h, p)

    def upload(self, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64)
        p = self.alloc(a.size)
        self.dev.call("spirk_copy_h2d", self.h, p, a.ctypes.data_as(C.c_void_p), a.size)
        return p

    def download(self, p, shape):
        out = np.empty(shape, dtype=np.float64)
        self.dev.call("spirk_copy_d2h", self.h, out.ctypes.data_as(C.c_void_p), p, out.size)
        return out

    def sync(self):
        self.dev.call("spirk_ctx_sync", self.h)

    def launches(self):
        return int(self.dev.lib.spirk_ctx_launch_count(self.h))

    def call(self, name, *args):
        self.dev.call(name, self.h, *args)

    def scalar_call(self, name, *args):
        r = C.c_double()
        self.dev.call(name, self.h, *args, C.byref(r))
        return r.value


def darr(values):
    a = np.ascontiguousarray(values, dtype=np.float64)
    return a.ctypes.data_as(dp), a
