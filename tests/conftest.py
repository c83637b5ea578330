import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG = 'bipartite-link-prediction_b200'


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pkg(sub=None):
    return importlib.import_module(PKG if sub is None else PKG + '.' + sub)


@pytest.fixture(scope='session')
def blp():
    return pkg()


@pytest.fixture(scope='session')
def built_lib():
    """The in-tree CUDA extension, compiled if stale (nvcc cross-compiles without a GPU)."""
    lib = pkg('_lib')
    lib.build()
    return lib.load()
