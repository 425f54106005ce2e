import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="session")
def ref_modules():
    """The reference's own kernels (oracle/_ref, built by oracle/build_ref.py); {} when absent."""
    from oracle import build_ref
    mods = {}
    for name in build_ref.MODULES:
        m = build_ref.load_ref(name)
        if m is not None:
            mods[name] = m
    return mods
