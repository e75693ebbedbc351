import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without a CUDA device (or without the built library) skips the gpu-marked tests
    instead of failing in DeviceFilter's constructor; `-m gpu` on the B200 box runs them."""
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    have_lib = os.path.exists(os.path.join(ROOT, "fast_slam_b200", "libfs2.so"))
    if have_gpu and have_lib:
        return
    why = "no CUDA device" if not have_gpu else "fast_slam_b200/libfs2.so is not built"
    skip = pytest.mark.skip(reason="needs a B200: " + why)
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
