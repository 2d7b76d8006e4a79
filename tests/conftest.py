import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


@pytest.fixture(scope="session")
def wva():
    import wva_import
    return wva_import.load()


@pytest.fixture(scope="session")
def oracle():
    import oracle as o
    o.build()
    o.lib()
    return o


@pytest.fixture(scope="session")
def ctx(wva):
    """One CUDA context for the whole GPU session (creation is the slow call)."""
    from inferno_autoscaler_b200 import binding
    c = binding.Context(0)
    yield c
    c.close()
