import importlib
import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def cu():
    """the product package (directory name has a hyphen)"""
    return importlib.import_module("chaos-ultra_b200")


@pytest.fixture(scope="session")
def provider(cu):
    if not _has_gpu():
        pytest.skip("no CUDA device")
    p = cu.CudaFractalRendererProvider()
    yield p
    p.close()
