import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def lib():
    """The CUDA library.  Built in-tree if missing; never replaced by a CPU path."""
    import adtomo_jl_b200 as A
    if not os.path.exists(A.LIB_PATH):
        A.build_library()
    return A


@pytest.fixture(scope="session")
def ctx(lib):
    c = lib.Context(0)
    yield c
    c.close()
