import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running CPU oracle case")


def pytest_collection_modifyitems(config, items):
    # a plain `pytest` on a box without a GPU skips the parity tests instead of failing them with
    # ERR_NO_DEVICE (the driver selects with -m gpu / -m "not gpu"; this is for everybody else)
    if has_cuda():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200): diffurch_b200 has no CPU execution path")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def dfb():
    import diffurch_b200
    return diffurch_b200


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
