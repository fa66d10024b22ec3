import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def wb():
    import worldb200
    return worldb200


@pytest.fixture(scope="session")
def signals():
    import worldb200  # noqa: F401  (registers the package)
    from worldb200 import signals as s
    return s
