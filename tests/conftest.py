import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def cuda_lib():
    """Loads libsoswsod_b200.so; GPU tests must fail loudly (not skip) when it is missing."""
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from sos_wsod_b200 import _lib

    return _lib.load()
