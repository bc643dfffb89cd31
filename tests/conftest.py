import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
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
def hostsim():
    """TEST-ONLY CPU build of the kernels' host/device logic headers (tests/hostsim)."""
    out_dir = os.path.join(ROOT, "tests", "_hostsim")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libhostsim.so")
    srcs = [os.path.join(ROOT, "tests", "hostsim", f) for f in sorted(os.listdir(os.path.join(ROOT, "tests", "hostsim"))) if f.endswith(".cpp")]
    hdr_dir = os.path.join(ROOT, "hevcbitstream_b200", "csrc")
    deps = srcs + [os.path.join(hdr_dir, f) for f in os.listdir(hdr_dir) if f.endswith(".h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I" + hdr_dir, "-o", so] + srcs)
    import ctypes

    return ctypes.CDLL(so)


@pytest.fixture(scope="session")
def ctx():
    import hevcbitstream_b200 as hb

    c = hb.Context(0)
    yield c
    c.close()
