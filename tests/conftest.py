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


@pytest.fixture(scope="session")
def ds_paper():
    """Paper geometry (DS.m:42-46: fs = 2.94 MHz, 2 subframes -> N = 7350, K = 1440 / 672, 32 pilots, 6 taps) with two
    SNR points (20 dB and the 32 dB of png/Figure5.png) to keep the oracle setup around a minute."""
    from oracle.ds import DSConfig, ds_setup
    return ds_setup(DSConfig.paper(M_SNR_dB=(20, 32)))
