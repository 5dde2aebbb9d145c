import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _native_libraries():
    """Build the host library and the oracle once; the CUDA library is cross-compiled by nvcc (no GPU needed)."""
    from epidemicsimulator_b200 import build
    build.build_host()
    build.build_cuda()
    from oracle import oracle_py
    oracle_py.build()
    yield


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
