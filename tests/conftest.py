import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # A fresh checkout has no built artefacts (they are git-ignored).  The TEST SESSION builds what it is about to test —
    # the library with nvcc for sm_100a (cross-compiles without a GPU), the oracle with g++ — exactly as
    # __graft_entry__.build() does; the product itself still refuses to run unbuilt (pyconnect6/_capi.py: product()).
    import c6lib
    from pyconnect6 import _capi
    if "C6_LIB" not in os.environ and not os.path.exists(_capi.LIB_PATH):
        c6lib.build_product()
    if not os.path.exists(c6lib.ORACLE_PATH) or not os.path.exists(c6lib.ORACLE_TESTS):
        c6lib.build_oracle()


@pytest.fixture(scope="session")
def oracle():
    import c6lib
    return c6lib.oracle()


@pytest.fixture(scope="session")
def lib():
    import c6lib
    return c6lib.product()
