import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def ref():
    """The CPU oracle (oracle/libbz2ref.so): the checker, never the thing under test."""
    from oracle import pyref
    pyref.build()
    return pyref


@pytest.fixture(scope="session")
def engine():
    """One GPU context of the product library.  Fails loudly if the CUDA extension is missing."""
    import bzip2_rust_b200 as bz
    eng = bz.Engine()
    yield eng
    eng.close()
