import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle (oracle/liboracle.so) -- the checker, never the thing under test."""
    from oracle import pyoracle

    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def rdf():
    import __graft_entry__ as entry

    if entry.needs_build():
        entry.build()
    import rust_dataframe_b200

    return rust_dataframe_b200


@pytest.fixture(scope="session")
def ctx(rdf):
    """One context for the whole session (GPU tests only)."""
    c = rdf.default_context()
    yield c
