import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def oracle():
    from oracle.orc import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle.orc import Ref
    if not Ref.available():
        pytest.skip("oracle/_ref/libdwt_ref.so not built (needs /root/reference)")
    return Ref()


@pytest.fixture(scope="session")
def dev():
    """The CUDA library, initialised.  Fails loudly (no skip, no fallback) when it cannot run."""
    import libdwt_b200 as d
    L = d.lib()
    L.init()
    return d
