import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ds_default():
    """Oracle setup of DoublySelectiveChannelEstimation.m with its default parameters."""
    from oracle.ds import DSConfig, ds_setup
    return ds_setup(DSConfig())


@pytest.fixture(scope="session")
def gpu_ctx(ds_default):
    """A device context fed with the oracle's setup outputs (parity inputs)."""
    from tests.helpers import context_from_oracle
    ctx = context_from_oracle(ds_default, max_batch=32)
    yield ctx
    ctx.close()
