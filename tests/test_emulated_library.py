"""The single-rank GPU parity suites, executed on the CPU against the WHOLE library compiled for the host
(tests/emu/build_context.py: csrc/context.cu, shell_force.cu, spectral.cu and every kernel instantiation, with the kernel
launches rewritten onto the CUDA execution-model emulator and the CUDA runtime / cuFFT replaced by host stand-ins).

`MLBM_EMULATED=1 pytest -m gpu` re-points the ctypes binding (tests/conftest.py) and runs the very same test functions the
B200 box runs -- golden vectors of the reference, oracle parity for every lattice x collision x scheme, the spectral forces and
spectra, the C++ template layer's binaries, and the multi-GPU suites with one PROCESS per rank (NCCL and the CUDA IPC mappings of
the direct peer halos emulated over POSIX shared memory, tests/emu/include/nccl.h, cuda_host_emu.h) -- except three full-size
cases.  The CPU suite runs all single-rank cases and a handful of multi-rank ones (worlds 2, 4 and 8); the whole multi-rank
set takes about four minutes (`MLBM_EMULATED=1 pytest tests -m gpu -n 3 -k slabs`).  What this checks: the
library's host logic (allocation, pitched copies, launch assembly, the observables / spectral pipelines) and the kernels' logic,
end to end through the C-ABI.  What it cannot check: anything about the hardware (rounding of the device math library, memory
model, speed), which is why the `-m gpu` run on the box stays the parity gate."""
import os
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_single_rank_gpu_suites_pass_on_the_emulated_library(cuda_lib):
    sys.path.insert(0, str(ROOT / "tests" / "emu"))
    import build_context
    build_context.build()                      # once, before the workers start
    result = _emulated_pytest([str(ROOT / "tests")], ranks=1, workers=min(8, os.cpu_count() or 1))
    tail = result.stdout[-3000:] + result.stderr[-1500:]
    assert result.returncode == 0, tail
    summary = re.search(r"(\d+) passed", result.stdout)
    assert summary and int(summary.group(1)) >= 130, tail      # the whole single-rank suite ran, not a skipped shell of it
    assert "failed" not in result.stdout.splitlines()[-1], tail


def _emulated_pytest(targets, ranks, workers):
    command = [sys.executable, "-m", "pytest", *targets, "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider", "--timeout=600"]
    try:
        import xdist  # noqa: F401
        if workers > 1:
            command += ["-n", str(workers)]
    except ImportError:
        pass
    return subprocess.run(command, capture_output=True, text=True, timeout=1500, cwd=ROOT,
                          env={**os.environ, "MLBM_EMULATED": "1", "MLBM_EMULATED_RANKS": str(ranks), "OMP_NUM_THREADS": "1"})


MULTI_RANK_SAMPLE = [
    # the configuration of the driver's 8-GPU scaling bench: boundary kernel stores into the neighbours' halo planes
    "tests/test_multi_gpu.py::test_slabs_reproduce_the_single_rank_result[D3Q19-BGK-None-On-async-peer-8]",
    "tests/test_multi_gpu.py::test_slabs_reproduce_the_single_rank_result[D3Q27-ELBM-Guo-On-sync-peer-4]",
    "tests/test_multi_gpu.py::test_slabs_reproduce_the_single_rank_result[D2Q9-ELBM-ExactDifferenceMethod-Off-async-2]",
    "tests/test_spectral_forces_gpu.py::test_spectral_forces_on_slabs[Turbulent2D-4-nccl]",
]


def test_multi_rank_sample_passes_on_the_emulated_library(cuda_lib):
    """One process per rank; halo exchange by NCCL send/recv (shared-memory stand-in), overlapped with the bulk kernel, and by
    direct stores into the neighbours' halo planes through CUDA IPC mappings (shared-memory stand-in) with the flag handshake;
    distributed spectral transform and the all-reduced observables / force projections."""
    sys.path.insert(0, str(ROOT / "tests" / "emu"))
    import build_context
    build_context.build()
    result = _emulated_pytest([str(ROOT / t) for t in MULTI_RANK_SAMPLE], ranks=8, workers=4)
    tail = result.stdout[-3000:] + result.stderr[-1500:]
    assert result.returncode == 0, tail
    assert re.search(rf"{len(MULTI_RANK_SAMPLE)} passed", result.stdout), tail
