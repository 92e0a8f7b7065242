import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ctx():
    """One shared bt_ctx for the GPU tests (sized for the largest test problem)."""
    import botsort_b200 as bs
    c = bs.Context(max_tracks=2304, max_dets=2304, feat_dim=2048)
    yield c
    c.close()
