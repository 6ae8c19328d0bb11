import pathlib
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: CPU test that takes tens of seconds (still part of the default run)")


@pytest.fixture(scope="session")
def built_library():
    """Build (if stale) and return the path of libnxfx_b200.so."""
    from networks_fenicsx_b200 import _build

    return _build.build_library()
