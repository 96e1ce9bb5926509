import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _native_built():
    lib = os.path.join(ROOT, "raytrace_b200", "lib")
    return all(os.path.exists(os.path.join(lib, n)) for n in ("librt_b200.so", "librt_host.so"))


@pytest.fixture(scope="session", autouse=True)
def native_libs():
    """Build the in-tree native libraries once if they are missing (nvcc cross-compiles without a GPU)."""
    if not _native_built():
        import __graft_entry__
        __graft_entry__.build()
    yield


@pytest.fixture(scope="session")
def gpu_present():
    # No silent fallback: on a box that has NVIDIA device nodes every failure is a real failure.
    if not os.path.exists("/dev/nvidiactl") and not os.path.exists("/dev/nvidia0"):
        pytest.skip("no NVIDIA device on this box (gpu tests run on the B200 box)")
    return True
