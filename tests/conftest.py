import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """libdwt_b200.so + oracle built (the driver calls __graft_entry__.build() first; this covers a bare checkout)."""
    import dwt_b200
    from oracle import pyoracle
    if not os.path.exists(dwt_b200.LIB_PATH) or not os.path.exists(pyoracle.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return True


@pytest.fixture(scope="session")
def oracle(built):
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def pins():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "pins.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def pins_big():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "pins_big.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def codec(built):
    """The CUDA codec context.  No skip, no fallback: on a box without a usable GPU this fails loudly."""
    import dwt_b200
    c = dwt_b200.Codec(0)
    yield c
    c.close()
