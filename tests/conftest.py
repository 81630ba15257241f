import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle behind the same C ABI (prefix orc_) — the checker."""
    from oracle_backend import OracleBackend

    return OracleBackend()


@pytest.fixture(scope="session")
def cuda():
    """The product backend (CUDA library).  Never falls back."""
    import dbsp_b200.runtime as rt

    return rt.Runtime(0)
