import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def oracle_built():
    """Build the CPU oracles (C restatement always; oracle/_ref only where /root/reference exists)."""
    from oracle import pyoracle
    if not pyoracle.available("port") or (os.path.exists("/root/reference/src/ESDFMap.cpp") and not pyoracle.available("ref")):
        pyoracle.build()
    return pyoracle
