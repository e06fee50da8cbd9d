import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device and the built library: skip them (instead of failing) when either is missing, unless they
    were asked for explicitly with -m gpu (then a missing GPU / library must fail loudly: there is no fallback path)."""
    if "gpu" in (config.getoption("-m") or ""):
        return
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    from sin3dm_b200 import _lib
    if have and os.path.exists(_lib.LIB_PATH):
        return
    skip = pytest.mark.skip(reason="needs a CUDA device and sin3dm_b200/lib/libsin3dm_b200.so (run with -m gpu on the B200 box)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
