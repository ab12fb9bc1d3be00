import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')
    config.addinivalue_line('markers', 'reference: needs the reference tree at /root/reference (build container only)')


@pytest.fixture(scope='session')
def built_lib():
    """The in-tree shared library; built on demand (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as ge
    if not os.path.exists(ge.LIB):
        ge.build()
    return ge.LIB


@pytest.fixture(scope='session')
def oracle_c():
    from oracle import cdp
    cdp.build()
    return cdp
