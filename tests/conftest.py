import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """Make sure the C-ABI library exists (compiles it if this checkout has not been built yet)."""
    import __graft_entry__ as g
    if not os.path.isfile(g.LIB):
        g.build()
    from deepatlas_b200 import _lib
    return _lib.load()


@pytest.fixture(scope="session")
def cuda(built_lib):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
