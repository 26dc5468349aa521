import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _has_gpu() -> bool:
    try:
        from rpg_open_remode_b200 import device_count
        return device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # A gpu-marked test must never silently pass on a CPU-only box.
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device (gpu tests run on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def small_sequence():
    """160x120 synthetic sequence (frames rendered lazily)."""
    from rpg_open_remode_b200 import synth
    return synth.SyntheticSequence(160, 120, seed=0x5EED0001)


@pytest.fixture(scope="session")
def qvga_sequence():
    from rpg_open_remode_b200 import synth
    return synth.SyntheticSequence(320, 240, seed=0x5EED0001)
