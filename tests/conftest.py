import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


EMULATED = os.environ.get("MLBM_EMULATED") == "1"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    if EMULATED:
        # TEST INFRASTRUCTURE: `MLBM_EMULATED=1 pytest -m gpu` runs the single-rank GPU suites on the CPU against the whole
        # library compiled for the host (tests/emu/build_context.py) -- a check of the library's host logic and of the kernels'
        # logic, never of hardware behaviour.  The binding is re-pointed from here; the product knows nothing about it.
        sys.path.insert(0, str(ROOT / "tests" / "emu"))
        import build_context
        if not os.environ.get("PYTEST_XDIST_WORKER"):
            for stale in Path("/dev/shm").glob("mlbm_emu_*"):     # segments of emulated ranks that were killed
                try:
                    stale.unlink()
                except OSError:
                    pass
        from metalbm_b200 import capi
        library = build_context.build(tuple(os.environ.get("MLBM_EMULATED_FLAGS", "").split()))   # e.g. -DMLBM_COLUMN_REGISTERS_Q27=1
        capi._library = capi.load_library(library)
        shim_dir = library.parent / "shim_lib"
        shim_dir.mkdir(exist_ok=True)
        link = shim_dir / "libmetalbm_b200.so"
        if link.is_symlink() or link.exists():
            link.unlink()
        link.symlink_to(library)
        os.environ["MLBM_SHIM_LIBDIR"] = str(shim_dir)
        import torch
        # emulated devices: one process per rank, NCCL and CUDA IPC over shared memory (MLBM_EMULATED_RANKS=1: single-rank cases only)
        ranks = int(os.environ.get("MLBM_EMULATED_RANKS", "8"))
        torch.cuda.device_count = lambda: ranks
        os.environ["MLBM_EMULATED_LIBRARY"] = str(library)   # for the rank processes the multi-GPU tests start


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
    if "test_spectral_forces_gpu" in node or "test_wide_lattices_gpu" in node or "test_bench_support_gpu" in node:
        return 2
    lowered = node.lower()
    if any(word in lowered for word in ("elbm", "entropic", "alpha", "logarithm")):
        return 1
    return 0


# full-size cases that only a GPU finishes: skipped when the suites run on the emulated library
EMULATION_TOO_LARGE = ("test_entropic_blocks_walking_several_planes", "test_mass_conservation_at_256_cubed",
                       "test_entropic_mass_conservation_at_baseline_sizes")


def pytest_collection_modifyitems(config, items):
    if EMULATED:
        skip = pytest.mark.skip(reason="full-size case: needs the GPU (the emulated library runs one CUDA thread at a time)")
        for item in items:
            if any(name in item.nodeid for name in EMULATION_TOO_LARGE):
                item.add_marker(skip)
    _order_gpu_items(items)


def _order_gpu_items(items):
    """GPU runs stop at the first failure (-x): keep the order within each group, but run the groups oldest code first, so
    that a defect in a newer kernel cannot hide the verdict on the older ones."""
    gpu = [item for item in items if item.get_closest_marker("gpu")]
    if not gpu:
        return
    ordered = iter(sorted(gpu, key=_gpu_order))   # stable
    items[:] = [next(ordered) if item.get_closest_marker("gpu") else item for item in items]


# ---- a multi-GPU box must not skip what it can run ---------------------------------------------------------------------
# Every multi-rank case skips itself with "needs N GPUs" on a smaller box.  The outcomes are recorded here and
# tests/test_zz_gpu_coverage.py (collected last) FAILS when a box with N or more GPUs skipped such a case, or when a box with
# two or more GPUs ran no multi-rank case at all: a parity run that silently degraded to one GPU is an error, not a skip.
GPU_OUTCOMES = {"skipped_needing": [], "ran": [], "multi_rank_ran": 0}


def pytest_runtest_logreport(report):
    import re
    if report.when == "setup" and report.skipped or report.when == "call" and report.skipped:
        text = str(report.longrepr)
        match = re.search(r"needs (\d+) GPUs", text)
        if match:
            GPU_OUTCOMES["skipped_needing"].append((report.nodeid, int(match.group(1))))
    elif report.when == "call" and report.passed:
        GPU_OUTCOMES["ran"].append(report.nodeid)
        # the multi-rank cases carry their world size in the test id ("[...-2]", "world2", "-2-" ...): counted by the tests themselves
