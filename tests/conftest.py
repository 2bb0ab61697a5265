import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu_lib():
    """libpinc_b200.so on a box with a GPU.  No fallback: a missing library or device fails the test."""
    from pinc_b200 import lib
    assert os.path.exists(lib.SO_PATH), "libpinc_b200.so missing (run __graft_entry__.build())"
    assert _has_gpu(), "no CUDA device visible"
    return lib.load()
