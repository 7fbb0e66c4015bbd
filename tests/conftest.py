import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (runs on the B200 box)")


@pytest.fixture(scope="session")
def port():
    from oracle.oracle import Port
    return Port()


@pytest.fixture(scope="session")
def ref():
    """The reference's own liblz4 (oracle/_ref); tests needing it are skipped when it was not built."""
    from oracle.oracle import Ref
    if not Ref.available():
        pytest.skip("oracle/_ref/libreflz4.so not built")
    return Ref()


@pytest.fixture(scope="session")
def codec(port):
    """Strongest checker available: compiled reference if present, else the pinned port."""
    from oracle.oracle import Ref
    return Ref() if Ref.available() else port


@pytest.fixture(scope="session")
def gpu():
    import plz4_b200 as P
    P.init(0)
    return P
