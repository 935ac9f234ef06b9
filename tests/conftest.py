import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The product library must exist before any test imports the package."""
    import __graft_entry__
    lib = os.path.join(ROOT, "vulkpy_b200", "libvulkpy_b200.so")
    if not os.path.exists(lib):
        __graft_entry__.build()
    from oracle import cpu_ref
    cpu_ref.build()


@pytest.fixture(scope="session")
def gpu():
    import vulkpy_b200 as vk
    return vk.GPU(0)


@pytest.fixture()
def rs():
    import numpy as np
    return np.random.default_rng(12345)
