import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run with -m gpu on a B200)')


@pytest.fixture(scope='session', autouse=True)
def _built():
    """Make sure the oracle libraries and libsbc.so exist (compiling needs no GPU)."""
    from oracle import pyorc
    if not os.path.exists(pyorc.PORT_PATH) or (os.path.isdir('/root/reference/src') and not pyorc.have_reference()):
        pyorc.build()
    from sde_burst_coast_b200 import swarm
    if not os.path.exists(swarm.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    yield
