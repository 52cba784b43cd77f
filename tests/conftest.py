import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


@pytest.fixture(scope="session")
def built():
    """The C-ABI library and the oracle, built in-tree (no-op when already built)."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def code576(built):
    from ldpcgputegra_b200 import Code
    return Code.load("576x288")
