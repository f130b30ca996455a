import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (oracle/bsk_oracle.c) -- the checker, never the thing under test."""
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def hostcore():
    from tests import hostcore_binding
    hostcore_binding.lib()
    return hostcore_binding


@pytest.fixture(scope="session")
def bsk():
    """The product package with a live CUDA device; GPU tests fail loudly if the extension is missing."""
    import torch
    import basilisk_env_b200 as b
    from basilisk_env_b200 import _native
    assert torch.cuda.is_available(), "GPU test without a CUDA device"
    _native.lib()
    return b
