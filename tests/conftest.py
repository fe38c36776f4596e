import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Builds (if needed) the product library and the oracle; returns nothing."""
    import __graft_entry__ as g
    g.build()


@pytest.fixture(scope="session")
def ctx(built):
    from baofit_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def golden_tree(tmp_path_factory):
    """A directory laid out like the reference checkout (config/ data/ demo/ models/) with the
    gzip-compressed covariance expanded, so that the .ini recipes run unmodified."""
    from baofit_b200 import workloads as W
    return W.make_golden_tree(str(tmp_path_factory.mktemp("golden")))
