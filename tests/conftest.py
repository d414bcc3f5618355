import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def fc():
    """the product package; building is __graft_entry__.build()'s job, loading fails loudly"""
    import fourierconvolutioncudalib_b200 as pkg
    pkg._lib.load()
    return pkg


@pytest.fixture(scope="session")
def dev(fc):
    d = fc.selectDeviceWithHighestComputeCapability()
    assert d >= 0, "no CUDA device"
    return d


@pytest.fixture(scope="session")
def reflib():
    import reflib as r
    return r.load()
