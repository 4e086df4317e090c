import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the live Python reference under /root/reference")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device: on a box without one they are skipped, not failed, so a plain `pytest tests`
    shows the CPU suite's real result."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_lib():
    """Host build of the scalar restatement (test infrastructure; built by __graft_entry__.build())."""
    from tests import _util
    return _util.oracle_lib()
