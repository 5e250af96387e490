import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def cuda_lib():
    """The built C-ABI library on a CUDA device; GPU tests fail loudly if it is missing."""
    import torch
    import fidm_b200  # noqa: F401
    from fidm_b200 import _lib
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    lib = _lib.lib()
    _lib.check(lib.fidm_device_supported(0), "device")
    return lib
