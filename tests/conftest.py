import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def oracle_lib():
    """Builds oracle/libchili_oracle.so (CPU restatement of chili; test infrastructure)."""
    from oracle import chili_oracle

    chili_oracle.build()
    return chili_oracle


@pytest.fixture(scope="session")
def crb_lib():
    """The product's C-ABI library; the gpu tests fail loudly if it has not been built."""
    from cellular_raza_b200 import _abi

    return _abi.load_library()
