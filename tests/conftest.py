import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module("camera-intrinsic-calibration-rs_b200")


@pytest.fixture(scope="session")
def oracle():
    import oracle as O  # oracle/oracle.py — test infrastructure
    O.lib()
    return O


def have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
