import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope='session')
def engine_factory():
    """Creates libncb engines on cuda:0; fails loudly when the library or the GPU is missing."""
    import torch
    from nanocompore_b200.engine import Engine
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    made = []

    def make(**kw):
        e = Engine(0, **kw)
        made.append(e)
        return e
    yield make
    for e in made:
        e.close()
