"""In-tree build of ``libmetalbm_b200.so`` (hand-written CUDA for sm_100a + the C-ABI).

``python -m metalbm_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles without a GPU;
the shared library stays next to the package so it travels to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PACKAGE_DIR = Path(__file__).resolve().parent
CSRC = PACKAGE_DIR / "csrc"
# experiments: MLBM_VARIANT=name MLBM_EXTRA_FLAGS="-DX=1 ..." builds libmetalbm_b200_name.so next to the product library
VARIANT = os.environ.get("MLBM_VARIANT", "")
# variant objects live outside the tree: only their .so travels to the GPU box (the snapshot is capped at 512 MiB)
BUILD = (Path("/tmp") / ("mlbm_build_" + VARIANT)) if VARIANT else PACKAGE_DIR / "_build"
LIBRARY = PACKAGE_DIR / ("libmetalbm_b200" + ("_" + VARIANT if VARIANT else "") + ".so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
         *os.environ.get("MLBM_EXTRA_FLAGS", "").split()]


def _stale(target: Path, sources: list) -> bool:
    if not target.is_file():
        return True
    stamp = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > stamp for s in sources)


def _compile(source: Path, headers: list) -> Path:
    obj = BUILD / (source.stem + ".o")
    if _stale(obj, [source, *headers]):
        cmd = [NVCC, *ARCH, *FLAGS, "-c", str(source), "-o", str(obj)]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        (BUILD / (source.stem + ".ptxas.log")).write_text(proc.stderr)
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed on {source.name}:\n{proc.stderr[-6000:]}")
    return obj


def build(verbose: bool = False) -> Path:
    BUILD.mkdir(exist_ok=True)
    sources = sorted(CSRC.glob("*.cu"))
    headers = sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [PACKAGE_DIR.parent / "include" / "metalbm_b200.h"]
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as pool:
        objects = list(pool.map(lambda s: _compile(s, headers), sources))
    if _stale(LIBRARY, objects):
        cmd = [NVCC, *ARCH, "-shared", "-o", str(LIBRARY), *map(str, objects), "-ldl"]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError("link failed:\n" + proc.stderr[-4000:])
    if verbose:
        print(LIBRARY)
    return LIBRARY


if __name__ == "__main__":
    build(verbose=True)
    sys.exit(0)
