import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (run with -m gpu on the GPU box)')


@pytest.fixture(scope='session')
def built():
    """Everything native is built once per session (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    lib = os.path.join(ROOT, 'bundle-adjustment_b200', 'libjaicov_b200.so')
    if not os.path.exists(lib) or not os.path.exists(os.path.join(ROOT, 'tests', '_build', 'libemul.so')) \
            or not os.path.exists(os.path.join(ROOT, 'bundle-adjustment_b200', 'libjaicov_host.so')):
        g.build()
    from oracle import oracle
    oracle.build_lib()
    return True
