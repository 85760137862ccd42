import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"
INPUTS = GOLDEN / "inputs"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def gpu():
    """The product library, initialised on cuda:0.  Fails (does not skip) if the extension is missing on a GPU box."""
    import imageencoder_b200 as ie
    from imageencoder_b200 import _lib
    _lib.check(ie.lib().ie_init(0))
    return ie
