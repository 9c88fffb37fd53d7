import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    """Fixtures recorded from the real reference by tests/golden/make_golden.py."""
    return np.load(os.path.join(ROOT, "tests", "golden", "golden.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O

    O.port()  # builds oracle/libporeover_oracle.so on first use
    return O
