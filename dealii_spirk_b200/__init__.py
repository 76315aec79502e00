"""dealii_spirk_b200 — B200-native implementation of the dealii-spirk hot path.

The package is a thin Python shell (ctypes loaders, build recipe, launcher) around two in-tree
shared libraries:
  libspirk_b200.so  — hand-written sm_100a CUDA kernels behind the C ABI include/spirk_b200.h
  libspirk_host.so  — the C++ host layer mirroring the reference's operator / preconditioner /
                      time-integrator API (include/spirk_host.h), linked against the former.
There is no CPU fallback: loading fails loudly if the CUDA library is missing, and every entry
point fails with SPIRK_ERR_DEVICE without a GPU.
"""
import os

from . import capi

HERE = os.path.dirname(os.path.abspath(__file__))
DEVICE_LIB_PATH = os.path.join(HERE, "libspirk_b200.so")
HOST_LIB_PATH = os.path.join(HERE, "libspirk_host.so")
TABLES_PATH = os.path.join(HERE, "tables", "butcher_tables.txt")

_device = None


def _point_at_torch_nccl():
    """Let the CUDA library dlopen the same NCCL build PyTorch bundles (see csrc/nccl_dl.h)."""
    if os.environ.get("SPIRK_NCCL_LIB"):
        return
    import importlib.util
    spec = importlib.util.find_spec("nvidia.nccl")
    if spec and spec.submodule_search_locations:
        cand = os.path.join(list(spec.submodule_search_locations)[0], "lib", "libnccl.so.2")
        if os.path.exists(cand):
            os.environ["SPIRK_NCCL_LIB"] = cand


def device_lib():
    """The product CUDA library (never the CPU oracle)."""
    global _device
    if _device is None:
        _point_at_torch_nccl()
        lib = capi.DeviceLib(DEVICE_LIB_PATH)
        if lib.backend() != "cuda-sm_100a":
            raise capi.SpirkError(f"{DEVICE_LIB_PATH} is not the CUDA build (backend={lib.backend()})")
        _device = lib
    return _device
