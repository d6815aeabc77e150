import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden_dir():
    return os.path.join(ROOT, 'tests', 'golden')


def pytest_sessionstart(session):
    """A fresh checkout has no in-tree library (*.so is git-ignored): build it once instead of failing every test.
    An existing library is left alone (the GPU box receives the one built here)."""
    from stereospike_b200 import _lib, build
    if not os.path.isfile(_lib.LIB_PATH) and os.path.isfile(os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')):
        build.build()
