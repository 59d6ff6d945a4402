import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def pytest_sessionstart(session):
    """Build the kernel library if it is missing or stale (nvcc cross-compiles sm_100a without a GPU), so a
    fresh checkout can run the suite directly; tests still fail loudly if the build is impossible."""
    try:
        from g2vlm_b200 import _lib
        _lib.build()
    except Exception as e:  # reported by the tests that need the library
        print(f"[conftest] kernel library build failed: {e}")


@pytest.fixture(autouse=True)
def _seed_default_generators():
    """Tests that draw device tensors without a generator (torch.randn(..., device="cuda")) must not change from run to
    run: a check that sits at the edge of its tolerance would then fail once in a while on the driver's box."""
    import torch

    torch.manual_seed(20261019)
    yield
