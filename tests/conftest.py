import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    import orc as _orc  # oracle/orc.py — test infrastructure
    _orc.build()
    return _orc


@pytest.fixture(scope="session")
def rtb():
    import ray_tracer_archive_b200 as r
    return r


@pytest.fixture(scope="session")
def ctx(rtb):
    """One context for the whole GPU session; fails loudly if the CUDA library cannot run."""
    c = rtb.Context(int(os.environ.get("LOCAL_RANK", "0")))
    yield c
    c.close()
