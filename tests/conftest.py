import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))

    return load


@pytest.fixture(scope="session")
def ref():
    """The UNMODIFIED reference, only where /root/reference exists (build container)."""
    from oracle import ref_loader

    if not ref_loader.available():
        pytest.skip("reference tree not present (GPU box): covered by tests/golden fixtures")
    return ref_loader.load()
