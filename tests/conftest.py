import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def tuning():
    """Set developer knobs of the native library (b200_set_tuning) for one test; restored afterwards."""
    import icicle_b200 as ib
    touched = {}

    def _set(name, value):
        touched.setdefault(name, ib.get_tuning(name))
        ib.set_tuning(name, value)

    yield _set
    for name, old in touched.items():
        ib.set_tuning(name, old)
