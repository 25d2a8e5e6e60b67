import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def oracle_mod():
    """The CPU oracle (test infrastructure): built on demand."""
    from oracle import oracle_py as O
    O.build()
    return O


@pytest.fixture(scope="session")
def built_lib():
    """Path of the CUDA C-ABI library; built on demand (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    path = os.path.join(ROOT, "rdis_b200", "librdis_b200.so")
    if not os.path.exists(path):
        g.build()
    return path
