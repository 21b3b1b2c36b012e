import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def teo():
    """(lib, handle) on cuda:0 — fails loudly if the extension is missing or no GPU."""
    import ctypes as C

    import torch

    from teochat_b200 import lib as L
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    lib = L.load()
    h = C.c_void_p()
    L.check(lib.teo_create(0, C.byref(h)), "teo_create")
    yield lib, h
    lib.teo_destroy(h)
