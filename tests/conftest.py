import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def naive_lib():
    """The plain-C double-loop oracle (oracle/naive_oracle.c), built on demand."""
    import ctypes
    import subprocess
    odir = os.path.join(ROOT, "oracle")
    so = os.path.join(odir, "_build", "libnaive_oracle.so")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(odir, "naive_oracle.c")):
        subprocess.check_call(["make", "-C", odir], stdout=subprocess.DEVNULL)
    return ctypes.CDLL(so)
