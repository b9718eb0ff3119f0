import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def native_lib():
    """The C-ABI library, built on demand (nvcc cross-compiles without a GPU)."""
    sys.path.insert(0, os.path.join(ROOT, "liteattention_b200", "csrc"))
    import build as la_build
    la_build.build()
    from liteattention_b200 import _native
    return _native
