import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session', autouse=True)
def _build_oracle():
    # the oracle is the checker for every test; building it is cheap (gcc, ~1 s)
    from oracle import build
    build.build_port()


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device in this container (GPU tests run under gpurun)')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)
