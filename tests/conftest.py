import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the live Python reference under /root/reference")


@pytest.fixture(scope="session")
def oracle_lib():
    """Host build of the scalar restatement (test infrastructure; built by __graft_entry__.build())."""
    from tests import _util
    return _util.oracle_lib()
