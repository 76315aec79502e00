"""ctypes binding of include/spirk_host.h (the C++ host layer that mirrors the reference's API)."""
import ctypes as C
import json
import os

import numpy as np

from . import capi

dp = C.POINTER(C.c_double)


class HostLib:
    def __init__(self, path, tables_path=None):
        if not os.path.exists(path):
            raise capi.SpirkError(f"{path} not found — run `python -c 'import __graft_entry__ as g; g.build()'`")
        if tables_path:
            os.environ["SPIRK_TABLES"] = tables_path
        self.lib = lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
        lib.spirk_host_last_error.restype = C.c_char_p
        lib.spirk_host_backend.restype = C.c_char_p
        lib.spirk_host_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int,
                                          C.POINTER(C.c_void_p)]
        for name in ("destroy", "setup", "step", "finish", "run"):
            getattr(lib, "spirk_host_" + name).argtypes = [C.c_void_p]
        lib.spirk_host_finished.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        lib.spirk_host_step_host.argtypes = [C.c_void_p, C.c_void_p]
        lib.spirk_host_timer_begin.argtypes = [C.c_void_p]
        lib.spirk_host_timer_end.argtypes = [C.c_void_p, dp]
        lib.spirk_host_set_compute_errors.argtypes = [C.c_void_p, C.c_int]
        lib.spirk_host_get_scalar.argtypes = [C.c_void_p, C.c_char_p, dp]
        lib.spirk_host_get_array.argtypes = [C.c_void_p, C.c_char_p, dp, C.c_int, C.POINTER(C.c_int)]
        lib.spirk_host_get_solution.argtypes = [C.c_void_p, C.c_void_p]
        lib.spirk_host_table_text.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        lib.spirk_host_gmg.argtypes = [C.c_int] * 8 + [dp]

    def backend(self):
        return self.lib.spirk_host_backend().decode()

    GMG_COLUMNS = ("dim", "degree", "n_procs", "n_cells", "n_dofs", "L", "n_iterations", "time")

    def gmg(self, dim, degree, n_refinements, mode, n_components=8, n_repetitions=10, device=0, n_procs=1):
        """one row of the reference's gmg.cc benchmark table (mode 0..3, gmg.cc:342-382)"""
        v = (C.c_double * 8)()
        self.check(self.lib.spirk_host_gmg(dim, device, degree, n_refinements, mode, n_components, n_repetitions, n_procs, v),
                   "spirk_host_gmg")
        return dict(zip(self.GMG_COLUMNS, list(v)))

    def check(self, st, what):
        if st != 0:
            raise capi.SpirkError(f"{what} failed ({st}): {self.lib.spirk_host_last_error().decode()}")


class Run:
    """One HeatEquation::Problem (reference main.cc:3014-3603) configured like the reference's JSON files."""

    def __init__(self, host: HostLib, params: dict, dim=3, device=0, nccl_id=None, rank=0, world=1, verbose=False):
        self.host = host
        self.h = C.c_void_p()
        text = json.dumps(params).encode()
        host.check(host.lib.spirk_host_create(text, 0, dim, device, nccl_id, rank, world, int(verbose), C.byref(self.h)),
                   "spirk_host_create")

    def close(self):
        if self.h:
            self.host.lib.spirk_host_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _call(self, name, *args):
        self.host.check(getattr(self.host.lib, "spirk_host_" + name)(self.h, *args), name)

    def setup(self):
        self._call("setup")

    def step(self):
        self._call("step")

    def step_host(self, host_solution: np.ndarray):
        assert host_solution.dtype == np.float64 and host_solution.flags.c_contiguous
        self._call("step_host", host_solution.ctypes.data_as(C.c_void_p))

    def finish(self):
        self._call("finish")

    def run(self):
        self._call("run")

    def finished(self):
        f = C.c_int()
        self._call("finished", C.byref(f))
        return bool(f.value)

    def timer_begin(self):
        self._call("timer_begin")

    def timer_end(self):
        v = C.c_double()
        self._call("timer_end", C.byref(v))
        return v.value

    def set_compute_errors(self, on):
        self._call("set_compute_errors", int(on))

    def scalar(self, key):
        v = C.c_double()
        self._call("get_scalar", key.encode(), C.byref(v))
        return v.value

    def array(self, key, cap=4096):
        buf = np.zeros(cap)
        n = C.c_int()
        self._call("get_array", key.encode(), buf.ctypes.data_as(dp), cap, C.byref(n))
        return buf[:n.value].copy()

    def solution(self):
        """the solution entries this process owns: all of them, or (spatial partition) the z-slab starting at the lexicographic
        index scalar("first_owned")"""
        out = np.empty(int(self.scalar("n_dofs_owned")))
        self._call("get_solution", out.ctypes.data_as(C.c_void_p))
        return out

    def table_text(self):
        buf = C.create_string_buffer(1 << 16)
        self._call("table_text", buf, len(buf))
        return buf.value.decode()
