"""pytest configuration: markers, repo imports, shared fixtures.

  -m "not gpu" : oracle vs golden vectors, host logic, ABI surface, host emulation of the device code, gloo x2
  -m gpu       : parity of the CUDA path (through the C ABI) against the oracle and the committed goldens
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def _have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (oracle/oracle_py.py)."""
    return entry.load_oracle()


@pytest.fixture(scope="session")
def pkg():
    """The product package; the library is (re)built with nvcc when sources are newer."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("rt_build", os.path.join(entry.PKG_DIR, "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    b.build()
    return entry.load_package()


@pytest.fixture(scope="session")
def hostsim():
    """tests/hostsim/libhostsim.so: the product's __host__ __device__ code compiled for the CPU (test only)."""
    import ctypes as C
    d = os.path.join(ROOT, "tests", "hostsim")
    so, src = os.path.join(d, "libhostsim.so"), os.path.join(d, "hostsim.cpp")
    hdrs = [os.path.join(entry.PKG_DIR, "csrc", h) for h in ("rt_math.cuh", "rt_types.h", "rt_shade.cuh", "rt_trace.cuh", "rt_build.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in [src] + hdrs):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared",
                        "-I/usr/local/cuda/include", src, "-o", so], check=True)
    return C.CDLL(so)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
