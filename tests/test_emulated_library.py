"""The single-rank GPU parity suites, executed on the CPU against the WHOLE library compiled for the host
(tests/emu/build_context.py: csrc/context.cu, communication.cu, analysis.cu, shell_force.cu, spectral.cu and every kernel instantiation, with the kernel
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
    # bench.run_ours as the driver launches it on N > 1 GPUs (1024^3 strong-scaled headline on a toy box, e2e leg, secondary rows)
    "tests/test_bench_support_gpu.py::test_the_multi_gpu_headline_of_the_bench_on_several_ranks[2]",
    "tests/test_multi_gpu.py::test_checkpoint_written_on_slabs_restarts_on_one_rank[2]",
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


def test_bench_runs_end_to_end_on_the_emulated_library(monkeypatch, capsys, cuda_lib):
    """bench.run_ours -- headline, per-launch kernel timing, the host-buffer e2e leg, the secondary workloads -- with the real
    Algorithm class over the emulated library on an 8^3 cube (torch.cuda's device calls faked away): every C-ABI call the
    bench makes exists and behaves; one complete JSON line comes out."""
    import json
    import types

    import torch

    sys.path.insert(0, str(ROOT / "tests" / "emu"))
    import build_context
    from metalbm_b200 import capi
    import bench
    monkeypatch.setattr(capi, "_library", capi.load_library(build_context.build()))
    monkeypatch.setattr(torch.cuda, "set_device", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "mem_get_info", lambda *a: (10 ** 11, 10 ** 11))
    real_empty = torch.empty
    monkeypatch.setattr(torch, "empty", lambda *a, **k: real_empty(*a, **{key: v for key, v in k.items() if key not in ("pin_memory", "device")}))
    monkeypatch.setattr(bench, "cpu_baseline_leg", lambda: {"value": 50.0, "unit": "MLUPS", "cores": 4, "kind": "reference", "sample": "fake"})
    monkeypatch.setattr(bench, "EDGE", 8)
    for name, work in bench.WORKLOADS.items():
        monkeypatch.setitem(work, "shape", (8, 8, 8) if work["shape"][2] > 1 else (16, 12, 1))
        if work["store_every"]:
            monkeypatch.setitem(work, "store_every", 2)
    monkeypatch.setattr(bench, "ALSO_SINGLE", [(n, d, e, m, 3) for n, d, e, m, _ in bench.ALSO_SINGLE])
    args = types.SimpleNamespace(gpus=1, steps=4, warmup=3, impl="ours", edge=8, variant=0, workload="d3q19_bgk_256", dtype="f64",
                                 overlap="On", halo="peer", eps=None, store_every=None, no_e2e=False, no_cpu_baseline=False,
                                 also="auto", also_timeout=600)
    assert bench.run_ours(args) == 0
    lines = [l for l in capsys.readouterr().out.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["value"] > 0 and line["gpu_launches"] >= 4 and line["roofline"]["kernel_launches_timed"] == 4
    assert line["e2e"]["value"] > 0 and "error" not in line["e2e"] and line["e2e"]["last_energy"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 19 * 8 ** 3 * 8 / 4
    assert [e["name"] for e in line["also"]] == [e[0] for e in bench.ALSO_SINGLE]
    for entry in line["also"]:
        assert "error" not in entry and "skipped" not in entry and entry["value"] > 0, entry
        assert abs(entry["mass_per_node"] - 1.0) < 1e-3, entry
    entropic = [e for e in line["also"] if "elbm" in e["name"]]
    assert entropic and all(0.0 <= e["alpha_off_shortcut_fraction_at_end"] <= 1.0 for e in entropic)
