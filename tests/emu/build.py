"""tests/emu/build.py -- TEST INFRASTRUCTURE ONLY.

Builds ``tests/emu/_build/libstep_emu.so``: the source of the fused step kernels compiled for the HOST under the
CUDA-execution-model emulator ``cuda_emu.h`` so that the CPU test-suite can check the kernels' logic without a GPU
(``tests/test_kernel_logic.py``).  The kernel source is taken from ``metalbm_b200/csrc`` at build time; the only edit is
the redirection of the one ``extern __shared__`` declaration to the emulator's shared-memory buffer."""
from __future__ import annotations

import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
CSRC = ROOT / "metalbm_b200" / "csrc"
BUILD = HERE / "_build"
LIBRARY = BUILD / "libstep_emu.so"
DYNAMIC_SHARED = "extern __shared__ __align__(16) unsigned char dynamicShared[];"


def build(flags: tuple = ()) -> Path:
    """``flags``: extra -D switches (the experimental kernel variants); every set of flags is its own library."""
    BUILD.mkdir(exist_ok=True)
    LIBRARY = BUILD / ("libstep_emu" + "".join("_" + f.lstrip("-D").lower() for f in flags) + ".so")
    sources = [CSRC / "step_kernel.cuh", CSRC / "lattice.cuh", CSRC / "log_table.inc", HERE / "cuda_emu.h", HERE / "step_emu.cpp",
               HERE / "include" / "cuda_runtime.h", Path(__file__)]
    if LIBRARY.is_file() and all(s.stat().st_mtime <= LIBRARY.stat().st_mtime for s in sources):
        return LIBRARY
    kernel = (CSRC / "step_kernel.cuh").read_text()
    if kernel.count(DYNAMIC_SHARED) != 1:
        raise RuntimeError("step_kernel.cuh: expected exactly one dynamic shared-memory declaration")
    kernel = kernel.replace(DYNAMIC_SHARED, "unsigned char* const dynamicShared = cuda_emu::dynamicSharedBase();")
    (BUILD / "step_kernel_emu.cuh").write_text(kernel)
    # hidden visibility + -Bsymbolic: libmetalbm_b200.so (loaded RTLD_GLOBAL by the tests) exports host stubs with the very
    # same mangled kernel names; the emulator must bind to its own definitions
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-ffp-contract=off", "-fno-strict-aliasing", "-w",
           "-fvisibility=hidden", "-fvisibility-inlines-hidden", "-Wl,-Bsymbolic", *flags,
           f"-I{HERE / 'include'}", f"-I{BUILD}", f"-I{CSRC}", str(HERE / "step_emu.cpp"), "-o", str(LIBRARY)]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("emulator build failed:\n" + proc.stderr[-6000:])
    return LIBRARY


if __name__ == "__main__":
    print(build())
