import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the read-only reference checkout at /root/reference")


@pytest.fixture(autouse=True)
def _developer_switches_follow_the_environment():
    """The library reads its PHB_* switches once; tests that monkeypatch one call ``phb_reload_tuning()`` and this
    fixture - set up before, hence torn down after, monkeypatch - re-reads the restored environment afterwards."""
    yield
    from phylo_utils_b200 import _lib
    _lib.lib().phb_reload_tuning()


def _gpu_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isdir("/root/reference/phylo_utils")
    skip_ref = pytest.mark.skip(reason="reference checkout not mounted on this box")
    for item in items:
        if "reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)
