import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_sessionstart(session):
    """The shared library is a build artefact (git-ignored): compile it when a fresh checkout has none.
    An existing library is used as it is -- on the GPU box the prebuilt one travels with the snapshot."""
    if not os.path.exists(os.path.join(ROOT, "ccvm_b200", "libccvm_b200.so")):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
