import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as graft  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def rd():
    return graft.load_package()


@pytest.fixture(scope="session")
def orc():
    return graft.load_oracle()


@pytest.fixture(scope="session")
def ctx(rd):
    """Engine context on cuda:0 — fails loudly when the extension or the device is missing."""
    return rd.engine.default_context()
