import os
import sys

import pytest

# the oracle's OpenMP threads sleep between parallel regions instead of spinning: the suite mixes
# many tiny regions with subprocesses (gloo ranks, the host driver), and spinning threads cost
# 2-4x the CPU time on a shared machine (libgomp reads this when it is loaded)
os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import musoracle
    musoracle.lib()
    return musoracle
