import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def cpu_dev():
    """TEST-ONLY CPU double of the C ABI (oracle/cpu_abi.cc)."""
    import subprocess
    from dealii_spirk_b200 import capi
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "_build/libspirk_cpu.so"])
    return capi.DeviceLib(os.path.join(ROOT, "oracle", "_build", "libspirk_cpu.so"))


@pytest.fixture(scope="session")
def gpu_dev():
    import dealii_spirk_b200
    return dealii_spirk_b200.device_lib()
