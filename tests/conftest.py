import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import OracleEngine, build
    build()
    return OracleEngine()


@pytest.fixture(scope="session")
def gpu():
    """The product engine. Fails loudly (no CPU fallback) if the CUDA library or
    a device is missing."""
    from cook_b200.engine import GpuEngine
    eng = GpuEngine()
    yield eng
    eng.close()
