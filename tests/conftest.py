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


def _gpu_order(item) -> int:
    """0: the BGK paths (validated on hardware longest), 1: the entropic kernels, 2: the array-type / spectral forces and
    the multi-speed lattices."""
    node = item.nodeid
    if "test_spectral_forces_gpu" in node or "test_wide_lattices_gpu" in node:
        return 2
    lowered = node.lower()
    if any(word in lowered for word in ("elbm", "entropic", "alpha", "logarithm")):
        return 1
    return 0


def pytest_collection_modifyitems(config, items):
    """GPU runs stop at the first failure (-x): keep the order within each group, but run the groups oldest code first, so
    that a defect in a newer kernel cannot hide the verdict on the older ones."""
    gpu = [item for item in items if item.get_closest_marker("gpu")]
    if not gpu:
        return
    ordered = iter(sorted(gpu, key=_gpu_order))   # stable
    items[:] = [next(ordered) if item.get_closest_marker("gpu") else item for item in items]
