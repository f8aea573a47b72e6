import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def rb():
    import rabbit_transcoding_b200 as rb
    return rb


@pytest.fixture(scope="session")
def checker_backend():
    """the CPU checker: the unmodified reference when oracle/_ref was built, else the C restatement"""
    from oracle import checker
    if checker.have_reference():
        return checker.Reference()
    if checker.have_port():
        return checker.Port()
    pytest.skip("no CPU checker built (cd oracle && make port)")


@pytest.fixture(scope="session")
def codec(rb):
    c = rb.codec.PCCCodecB200(device=0)
    yield c
    c.close()
