import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


H2O = """
O       0.0000000000    -0.0000000000     0.1174000000
H      -0.7570000000    -0.0000000000    -0.4696000000
H       0.7570000000     0.0000000000    -0.4696000000
"""


@pytest.fixture(scope="session")
def h2o_atom():
    return H2O


@pytest.fixture(scope="session", autouse=True)
def _ensure_native_built():
    """The CUDA library and the oracle are build products (git-ignored): build them on demand so
    that a fresh checkout can run the suite (nvcc cross-compiles without a GPU, ~80 s)."""
    from joltqc_b200 import build as b
    if not os.path.exists(b.LIB):
        b.build()
    from oracle import oracle
    oracle.build()
