import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


@pytest.fixture(scope="session")
def cuda_lib():
    """The built C-ABI library (built here on first use; nvcc cross-compiles without a GPU)."""
    import spectrum_analyzer_b200 as sa
    sa.build()
    return sa.cabi.load()


@pytest.fixture(scope="session")
def gpu_ctx(cuda_lib):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import spectrum_analyzer_b200 as sa
    return sa.default_context(0)
