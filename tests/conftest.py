import importlib
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


def ensure_library():
    """librt_b200.so is a build artefact (git-ignored): build it when missing or older than its sources
    (nvcc cross-compiles for sm_100a without a GPU). The product itself never builds or falls back."""
    csrc = os.path.join(ROOT, "sycl-ray-tracer_b200", "csrc")
    lib = os.path.join(ROOT, "sycl-ray-tracer_b200", "librt_b200.so")
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".h"))] + [os.path.join(ROOT, "include", "rt_api.h")]
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(f) for f in srcs):
        subprocess.run(["make", "-s", "-C", csrc, "-j4"], check=True)


def load_package():
    ensure_library()
    return importlib.import_module("sycl-ray-tracer_b200")


@pytest.fixture(scope="session")
def pkg():
    return load_package()


@pytest.fixture(scope="session")
def scenes(pkg):
    return importlib.import_module("sycl-ray-tracer_b200.scenes")


@pytest.fixture(scope="session")
def oracle():
    import _oracle
    return _oracle


@pytest.fixture(scope="session")
def hostemu():
    import _hostemu
    return _hostemu


@pytest.fixture(scope="session")
def app(pkg):
    """one rt_context for the whole GPU session"""
    a = pkg.App(0)
    yield a
    a.close()
