"""Build recipes (in-tree, no JIT cache): the CUDA device library, the C++ host library, and
(test infrastructure) the CPU oracle libraries.  nvcc cross-compiles sm_100a without a GPU."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DEVICE_LIB = os.path.join(HERE, "libspirk_b200.so")
HOST_LIB = os.path.join(HERE, "libspirk_host.so")
NVCC = os.environ.get("SPIRK_NVCC", "/usr/local/cuda/bin/nvcc")
# the image exports CXX=/opt/gcc/bin/g++ (no libgomp.spec); use the distro compiler explicitly
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def device_sources():
    d = os.path.join(HERE, "csrc")
    return [os.path.join(d, f) for f in sorted(os.listdir(d))] + [os.path.join(ROOT, "include", "spirk_b200.h")]


def build_device(force=False, verbose_ptxas=False, defines=(), out=None, objdir=None):
    """One object per translation unit (the main unit + one unit per mode of the plane-streaming cell operator,
    csrc/v3_mode*.cu), compiled side by side, each single-threaded: nvcc --split-compile gave a different register
    allocation of the hot kernels from build to build (see profiles/README.md), separate units are deterministic."""
    srcs = device_sources()
    out = out or DEVICE_LIB
    if not force and not _newer(out, srcs):
        return out
    csrc = os.path.join(HERE, "csrc")
    units = ["spirk_b200.cu"] + sorted(f for f in os.listdir(csrc) if f.startswith("v3_mode") and f.endswith(".cu"))
    objdir = objdir or os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    # (the CUDA runtime is linked statically, nvcc's default: the library must not depend on which libcudart the
    # Python environment happens to put first on the loader path)
    base = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-ccbin", GXX,
            "-Xcompiler", "-fPIC,-O3"] + ["-D" + d for d in defines]
    if verbose_ptxas:
        base.insert(1, "-Xptxas=-v")
    procs, objs = [], []
    headers = [s_ for s_ in srcs if not s_.endswith(".cu")]
    for u in units:
        obj = os.path.join(objdir, u[:-3] + ".o")
        objs.append(obj)
        if not force and not _newer(obj, headers + [os.path.join(csrc, u)]):
            continue  # this unit is up to date (a unit depends on its own source and the headers)
        cmd = base + ["-c", "-o", obj, os.path.join(csrc, u)]
        print("+", " ".join(cmd), flush=True)
        procs.append((u, subprocess.Popen(cmd)))
    for u, p in procs:
        if p.wait() != 0:
            raise subprocess.CalledProcessError(p.returncode, u)
    _run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-ccbin", GXX, "-o", out] + objs + ["-ldl"])
    return out


def build_variant(tag, defines, force=False):
    """A kernel-experiment build of the device library with extra -D flags (SPIRK_V3_MINB / _NBUF / _NAC ...):
    dealii_spirk_b200/variants/libspirk_b200_<tag>.so, loaded by tools/bench_vmult.py --lib."""
    vdir = os.path.join(HERE, "variants")
    os.makedirs(vdir, exist_ok=True)
    return build_device(force, defines=defines, out=os.path.join(vdir, f"libspirk_b200_{tag}.so"),
                        objdir=os.path.join(vdir, "obj_" + tag))


def host_sources():
    d = os.path.join(HERE, "host")
    return [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith((".cc", ".h"))] + \
        [os.path.join(ROOT, "include", "spirk_b200.h"), os.path.join(ROOT, "include", "spirk_host.h")]


def build_host(force=False):
    srcs = [s for s in host_sources() if os.path.exists(s)]
    if not force and not _newer(HOST_LIB, srcs + [DEVICE_LIB]):
        return HOST_LIB
    inc = "-I" + os.path.join(ROOT, "include")
    _run([GXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-reorder", "-Wno-comment", inc, "-o", HOST_LIB,
          os.path.join(HERE, "host", "host_capi.cc"), "-L" + HERE, "-lspirk_b200", "-Wl,-rpath,$ORIGIN"])
    # stand-alone driver with the reference's command line (main.cc:3608-3791)
    _run([GXX, "-O2", "-std=c++17", inc, "-o", os.path.join(HERE, "spirk_main"), os.path.join(HERE, "host", "main.cc"),
          "-L" + HERE, "-lspirk_host", "-lspirk_b200", "-Wl,-rpath,$ORIGIN"])
    return HOST_LIB


def build_oracle():
    _run(["make", "-C", os.path.join(ROOT, "oracle"), "all"])


def build_all(force=False):
    build_device(force)
    build_host(force)
    build_oracle()


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
