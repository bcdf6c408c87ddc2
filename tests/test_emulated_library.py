"""The single-rank GPU parity suites, executed on the CPU against the WHOLE library compiled for the host
(tests/emu/build_context.py: csrc/context.cu, shell_force.cu, spectral.cu and every kernel instantiation, with the kernel
launches rewritten onto the CUDA execution-model emulator and the CUDA runtime / cuFFT replaced by host stand-ins).

`MLBM_EMULATED=1 pytest -m gpu` re-points the ctypes binding (tests/conftest.py) and runs the very same test functions the
B200 box runs -- golden vectors of the reference, oracle parity for every lattice x collision x scheme, the spectral forces and
spectra, the C++ template layer's binaries -- except the multi-rank ones and three full-size cases.  What this checks: the
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
    command = [sys.executable, "-m", "pytest", str(ROOT / "tests"), "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider", "--timeout=300"]
    try:
        import xdist  # noqa: F401
        command += ["-n", str(min(8, os.cpu_count() or 1))]
    except ImportError:
        pass
    result = subprocess.run(command, capture_output=True, text=True, timeout=1500, cwd=ROOT,
                            env={**os.environ, "MLBM_EMULATED": "1", "OMP_NUM_THREADS": "1"})
    tail = result.stdout[-3000:] + result.stderr[-1500:]
    assert result.returncode == 0, tail
    summary = re.search(r"(\d+) passed", result.stdout)
    assert summary and int(summary.group(1)) >= 120, tail      # the whole single-rank suite ran, not a skipped shell of it
    assert "failed" not in result.stdout.splitlines()[-1], tail
