import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def monkeypatch(monkeypatch):
    """monkeypatch whose setenv / delenv also make libcellcomm_b200 re-read its CC_* knobs (the
    library caches them at the first launch), and which re-reads them once more after the test's
    changes have been undone."""
    real_setenv, real_delenv = monkeypatch.setenv, monkeypatch.delenv

    def _reload():
        try:
            from cellcomm_b200 import _lib
            if _lib._LIB is not None:
                _lib._LIB.cc_reload_env()
        except Exception:
            pass

    def setenv(name, value, prepend=None):
        real_setenv(name, value, prepend)
        if name.startswith("CC_"):
            _reload()

    def delenv(name, raising=True):
        real_delenv(name, raising)
        if name.startswith("CC_"):
            _reload()

    monkeypatch.setenv, monkeypatch.delenv = setenv, delenv
    yield monkeypatch
    monkeypatch.undo()
    _reload()
