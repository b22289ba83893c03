import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def ctx():
    from orb_slam3_study_kr_b200 import api
    try:
        c = api.Context(0)
    except api.BagpuError as e:            # no CUDA device (or library not built): GPU tests skip, they never fall back
        pytest.skip(f"libbagpu needs a CUDA device: {e}")
    yield c
    c.close()


def pytest_collection_modifyitems(config, items):
    """Plain `pytest tests` on a host without a GPU: skip everything marked gpu instead of erroring."""
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
