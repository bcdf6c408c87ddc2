"""tests/emu/build_context.py -- TEST INFRASTRUCTURE ONLY.

Builds ``tests/emu/_build/libmetalbm_emu.so``: the WHOLE product library (csrc/context.cu, communication.cu, analysis.cu, shell_force.cu, spectral.cu and
the kernel instantiations) compiled for the HOST, so that the CPU test-suite can execute the library's host logic -- the
C-ABI entry points, allocation and pitched copies, the per-step orchestration, the observables and spectral pipelines --
around the emulated kernels (``cuda_emu.h``) without a GPU.  ``tests/test_emulated_library.py`` points the ctypes binding at
it and re-runs single-rank cases of the GPU parity suites.

The sources are taken from ``metalbm_b200/csrc`` at build time.  Edits, all mechanical:
  * ``kernel<<<grid, block[, shared[, stream]]>>>(args)``  ->  ``cuda_emu::launchKernel(kernel, grid, block, shared, args)``;
  * the one ``extern __shared__`` declaration of each file -> the emulator's shared-memory buffer;
  * ``asm volatile("trap;")`` -> ``abort()``;
  * spectral.cu resolves its cuFFT entry points from the naive host transforms of ``include/cufft.h`` instead of dlopen,
    communication.cu its NCCL entry points from the shared-memory NCCL of ``include/nccl.h`` (one process per rank).
``cuda_runtime.h``, ``nccl.h`` and ``cufft.h`` resolve to the stand-ins under ``tests/emu/include``: device memory is host
memory (poisoned at allocation), streams run immediately, there is one device and one rank.  Nothing under
``metalbm_b200/`` can reach this library; it says nothing about hardware behaviour or speed.

Variants: ``MLBM_EMULATED_FLAGS="-DMLBM_COLUMN_REGISTERS_Q27=1 -DMLBM_LOG_TABLE_SPLIT_Q9=1"`` builds the kernel variants of step_kernel.cuh;
``MLBM_EMULATED_FLAGS="-fsanitize=address -DMLBM_EMU_ASAN" LD_PRELOAD=$(gcc -print-file-name=libasan.so)
ASAN_OPTIONS=detect_leaks=0 MLBM_EMULATED=1 MLBM_EMULATED_RANKS=1 pytest tests -m gpu`` runs the single-rank suites with
"device" allocations on the sanitised heap, so that an index error of a kernel or of the host code lands in a red zone.
"""
from __future__ import annotations

import os
import re
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
CSRC = ROOT / "metalbm_b200" / "csrc"
BUILD = HERE / "_build" / "context"
LIBRARY = HERE / "_build" / "libmetalbm_emu.so"
DYNAMIC_SHARED = re.compile(r"extern __shared__ (?:__align__\(16\) )?(unsigned char|double) (\w+)\[\];")


def _matching(text: str, start: int, open_char: str, close_char: str) -> int:
    """Index just past the bracket that closes the one at ``start``."""
    depth = 0
    for i in range(start, len(text)):
        if text[i] == open_char:
            depth += 1
        elif text[i] == close_char:
            depth -= 1
            if depth == 0:
                return i + 1
    raise ValueError("unbalanced brackets")


def _split_arguments(text: str) -> list:
    parts, depth, current = [], 0, ""
    for ch in text:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(current.strip())
            current = ""
        else:
            current += ch
    parts.append(current.strip())
    return parts


def rewrite_launches(text: str) -> str:
    """``expr<<<config>>>(args)`` -> ``cuda_emu::launchKernel(expr, grid, block, shared, args)``."""
    out, position = "", 0
    while True:
        at = text.find("<<<", position)
        if at < 0:
            return out + text[position:]
        # the kernel expression: identifiers, ::, ->, and one trailing template argument list
        begin = at
        if text[begin - 1] == ">":
            depth, i = 0, begin - 1
            while True:
                if text[i] == ">":
                    depth += 1
                elif text[i] == "<":
                    depth -= 1
                    if depth == 0:
                        break
                i -= 1
            begin = i
        while begin > 0 and (text[begin - 1].isalnum() or text[begin - 1] in "_:" or text[begin - 2:begin] == "->" or text[begin - 1] == ">" and text[begin - 2] == "-"):
            begin -= 2 if text[begin - 2:begin] == "->" else 1
        kernel = text[begin:at]
        close = text.index(">>>", at)
        config = _split_arguments(text[at + 3:close])
        assert 2 <= len(config) <= 4, config
        shared = config[2] if len(config) > 2 else "0"
        paren = close + 3
        while text[paren].isspace():
            paren += 1
        assert text[paren] == "(", text[at - 40:paren + 10]
        end = _matching(text, paren, "(", ")")
        arguments = text[paren + 1:end - 1].strip()
        call = f"cuda_emu::launchKernel({kernel}, dim3({config[0]}), dim3({config[1]}), (size_t)({shared})" + (", " + arguments if arguments else "") + ")"
        out += text[position:begin] + call
        position = end


def transform(name: str) -> str:
    text = (CSRC / name).read_text()
    text = rewrite_launches(text)
    assert "<<<" not in text
    text = DYNAMIC_SHARED.sub(lambda m: f"{m.group(1)}* const {m.group(2)} = reinterpret_cast<{m.group(1)}*>(cuda_emu::dynamicSharedBase());", text)
    text = text.replace('asm volatile("trap;");', "abort();")
    text = text.replace('#include "../../include/metalbm_b200.h"', f'#include "{ROOT / "include" / "metalbm_b200.h"}"')
    if name == "communication.cu":
        start = text.index("const NcclApi* loadNccl(const char** error) {")
        end = _matching(text, text.index("{", start), "{", "}")
        text = (text[:start] + "const NcclApi* loadNccl(const char** error) {\n  (void)error;\n"
                "  static NcclApi api = {ncclGetUniqueId, ncclCommInitRank, ncclCommDestroy, ncclSend, ncclRecv, ncclGroupStart, ncclGroupEnd,\n"
                "                        ncclAllReduce, ncclGetErrorString};\n  return &api;\n}" + text[end:])
    if name == "spectral.cu":
        start = text.index("const CufftApi* loadCufft(std::string* error) {")
        end = _matching(text, text.index("{", start), "{", "}")
        text = (text[:start] + "const CufftApi* loadCufft(std::string* error) {\n  (void)error;\n"
                "  static CufftApi api = {cufftCreate, cufftMakePlanMany64, cufftSetStream, cufftExecD2Z, cufftExecZ2Z, cufftDestroy};\n"
                "  return &api;\n}" + text[end:])
    return text


def build(flags: tuple = ()) -> Path:
    """``flags``: extra -D switches (the experimental kernel variants); every set of flags is its own library."""
    global BUILD, LIBRARY
    suffix = "".join("_" + re.sub(r"[^a-z0-9]+", "", f.lower()) for f in flags)
    BUILD = HERE / "_build" / ("context" + suffix)
    LIBRARY = HERE / "_build" / ("libmetalbm_emu" + suffix + ".so")
    BUILD.mkdir(parents=True, exist_ok=True)
    sources = sorted(CSRC.glob("*.cu"))
    inputs = [*sources, *CSRC.glob("*.cuh"), *CSRC.glob("*.h"), *CSRC.glob("*.inc"), HERE / "cuda_emu.h", Path(__file__),
              *(HERE / "include").glob("*.h"), ROOT / "include" / "metalbm_b200.h"]
    if LIBRARY.is_file() and all(s.stat().st_mtime <= LIBRARY.stat().st_mtime for s in inputs):
        return LIBRARY
    for header in [*CSRC.glob("*.cuh"), *CSRC.glob("*.h"), *CSRC.glob("*.inc")]:
        (BUILD / header.name).write_text(transform(header.name) if header.suffix in (".cuh", ".h") else header.read_text())

    def compile_one(source: Path) -> Path:
        translated = BUILD / (source.stem + ".cpp")
        translated.write_text(transform(source.name))
        obj = BUILD / (source.stem + ".o")
        cmd = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-c", "-ffp-contract=off", "-fno-strict-aliasing", "-w", "-DMLBM_EMU_HOST", *flags,
               f"-I{HERE / 'include'}", f"-I{BUILD}", str(translated), "-o", str(obj)]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError(f"emulated build of {source.name} failed:\n" + proc.stderr[-6000:])
        return obj

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as pool:
        objects = list(pool.map(compile_one, sources))
    # -Bsymbolic: the tests load libmetalbm_b200.so as well, with the very same exported names
    cmd = ["g++", "-shared", "-Wl,-Bsymbolic", *[f for f in flags if f.startswith("-fsanitize")], "-o", str(LIBRARY), *map(str, objects),
           "-ldl", "-lrt", "-pthread"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("emulated link failed:\n" + proc.stderr[-4000:])
    return LIBRARY


if __name__ == "__main__":
    print(build())
