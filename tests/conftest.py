import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _env_options():
    """GMQL_* environment variables act as per-context options in the tests (gmql_b200/testing.py); the library itself never
    reads the environment."""
    try:
        from gmql_b200 import _lib as L
        from gmql_b200.testing import install_env_hooks
        install_env_hooks(L.load_library())
    except Exception:
        pass            # library not built: the tests that need it fail on their own
    yield


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
