import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the oracle (checker) and the native library exist; building them is not using them."""
    import oracle
    oracle.build()
    lib = os.path.join(ROOT, "fem_2d_b200", "libfem2d_b200.so")
    if not os.path.exists(lib):
        from fem_2d_b200 import build as _b  # noqa: F401  (import fails loudly if nvcc is missing)
    yield
