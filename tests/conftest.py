import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import oracle
    oracle.build()
    return oracle.lib()


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library; building it is part of the CPU-side check (nvcc cross-compiles)."""
    from metalbm_b200 import capi
    if not capi.LIBRARY_PATH.is_file():
        from metalbm_b200 import build
        build.build()
    return capi.load_library()
