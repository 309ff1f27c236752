import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def srk_ops():
    """The product operator layer; GPU tests fail loudly if the CUDA library is not built/loadable."""
    import torch
    assert torch.cuda.is_available(), "GPU test collected without a GPU"
    from ml_super_resolution_b200 import ops
    ops.handle()
    return ops
