import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure; see oracle/oracle.h)."""
    import oracle as o

    o.build()
    return o


@pytest.fixture(scope="session")
def native():
    """libragera.so, built in-tree. Fails loudly if it cannot be built/loaded — no fallback."""
    from rag_era_b200 import _native

    _native.build()
    return _native


@pytest.fixture(scope="session")
def golden():
    import json

    with open(os.path.join(ROOT, "tests", "golden", "kat_rrf.json")) as f:
        return json.load(f)
