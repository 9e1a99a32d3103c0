import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` through gpurun)")


@pytest.fixture(scope="session")
def native_lib():
    """Build (if needed) and load libcoddvec.so."""
    from codd_query_engine_b200 import build, native

    if not native.library_path().exists():
        build.build()
    return native.load_library()
