"""GPU checks of the C-ABI entry points that exist for bench.py's secondary workloads: the synthetic initial field made on the
device (mlbm_init_synthetic) and the asynchronous run loop with a chosen stored mode (mlbm_run_async_stored).  Newest code:
ordered last by tests/conftest.py."""
import numpy as np
import pytest

from helpers import relative_error
from metalbm_b200.algorithm import Algorithm
from metalbm_b200.capi import make_config
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("lattice,shape,equilibrium,dtype", [("D2Q9", (12, 10, 1), "TruncationMa3", "F64"), ("D2Q9", (9, 130, 1), "Exact", "F64"),
                                                             ("D3Q19", (6, 5, 7), "TruncationMa3", "F64"), ("D3Q27", (4, 6, 5), "Exact", "F32"),
                                                             ("D2Q21", (8, 6, 1), "TruncationMa3", "F64")])
def test_synthetic_field_made_on_the_device(lattice, shape, equilibrium, dtype):
    """mlbm_init_synthetic == initDistribution (Initialize.h:106-117) of the same density ripple and Taylor-Green-like velocity
    evaluated on the host (the field bench.py's headline workload uploads)."""
    cfg = make_config(lattice=lattice, shape=shape, equilibrium=equilibrium, dtype=dtype, tau=0.6)
    dim = 3 if shape[2] > 1 else 2
    x = (2 * np.pi * np.arange(shape[0]) / shape[0])[:, None, None]
    y = (2 * np.pi * np.arange(shape[1]) / shape[1])[None, :, None]
    z = (2 * np.pi * np.arange(shape[2]) / shape[2])[None, None, :]
    ones = np.ones(shape)
    density = 1.0 + 0.05 * np.sin(x) * np.cos(y) * np.cos(z) * ones
    if dim == 3:
        velocity = np.stack([0.04 * np.sin(x) * np.cos(y) * np.cos(z), -0.04 * np.cos(x) * np.sin(y) * np.cos(z),
                             0.02 * np.cos(x) * np.cos(y) * np.sin(z)])
    else:
        velocity = np.stack([0.04 * np.sin(y) * ones, 0.04 * np.cos(x) * ones])
    want = O.init_equilibrium(cfg, density, velocity)
    with Algorithm(cfg, host_fields=False) as algorithm:
        algorithm.init_synthetic(0.05, 0.04)
        algorithm.pack()
        got = algorithm.distribution.get_interior().astype(np.float64)
    assert relative_error(got, want) <= (1e-14 if dtype == "F64" else 1e-6)


def test_run_async_with_energy_only_stored_steps():
    """mlbm_run_async_stored with mode 2: energy / mass / Mach of the stored steps without field arrays (what a slab that fills
    the GPU can afford); the same numbers as mode 1, the enstrophy not available."""
    cfg = make_config(lattice="D3Q19", shape=(8, 6, 5), collision="BGK", forcing_scheme="Guo", force="Kolmogorov", tau=0.6,
                      amplitude=(1e-4, 1e-4, 1e-4), wavelength=(4.0, 4.0, 4.0))
    rows = {}
    for mode in (1, 2):
        with Algorithm(cfg, host_fields=False) as algorithm:
            algorithm.init_synthetic(0.05, 0.05)
            algorithm.run(1, 10, 5, stored_mode=mode)           # steps 1..10, stored on 5 and 10
            rows[mode] = algorithm.observables()
            algorithm.pack()
            rows[mode, "f"] = algorithm.distribution.get_interior()
    assert np.array_equal(rows[1, "f"], rows[2, "f"])
    assert rows[1][0] == rows[2][0] and rows[1][2] == rows[2][2] and rows[1][3] == rows[2][3]
    assert np.isfinite(rows[1][1]) and rows[1][1] > 0 and np.isnan(rows[2][1])
    with pytest.raises(Exception):
        with Algorithm(cfg) as algorithm:
            algorithm.run(1, 2, 1, stored_mode=3)


def test_secondary_workloads_of_the_bench_at_toy_sizes(monkeypatch):
    """bench.measure_also with the real context on every single-GPU secondary workload, the grids shrunk to toys: the calls it
    makes exist and agree with the library (device-side init, perturbation, stored modes, kernel timing, observables)."""
    import types

    import torch

    import bench
    if not torch.cuda.is_available():          # the emulated library: there is no device memory to ask about
        monkeypatch.setattr(torch.cuda, "mem_get_info", lambda *a: (10 ** 11, 10 ** 11))
    for name, work in bench.WORKLOADS.items():
        monkeypatch.setitem(work, "shape", (8, 6, 5) if work["shape"][2] > 1 else (12, 10, 1))
        if work["store_every"]:
            monkeypatch.setitem(work, "store_every", 2)
    args = types.SimpleNamespace(overlap="On", halo="peer", variant=0)
    identity = lambda v: v  # noqa: E731
    for entry in bench.ALSO_SINGLE:
        name, dtype, eps, mode, steps = entry
        out = bench.measure_also((name, dtype, eps, mode, 4), args, 0, 1, 0, lambda: None, identity, identity, 6500.0)
        assert "skipped" not in out and out["value"] > 0 and out["kernel_launches_timed"] == 4, out
        assert abs(out["mass_per_node"] - 1.0) < 1e-3 and out["energy"] > 0 and 0 < out["mach"] < 0.5, out
        if bench.WORKLOADS[name]["store_every"]:
            assert out["stored_step_ms"] is not None and out["stored_step_ms"] >= 0


def test_bench_prints_one_contract_line_on_a_small_cube():
    """`python bench.py --edge 32`: the whole driver on a real device at a size that takes a second; every key of the contract,
    the e2e leg through host buffers, the roofline of the fused kernel."""
    import json
    import subprocess
    import sys
    from pathlib import Path

    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device (bench.py drives torch.cuda directly)")
    root = Path(__file__).resolve().parent.parent
    result = subprocess.run([sys.executable, str(root / "bench.py"), "--edge", "32", "--steps", "20", "--warmup", "3", "--no-cpu-baseline"],
                            capture_output=True, text=True, timeout=600, cwd=root)
    assert result.returncode == 0, result.stderr[-3000:]
    lines = [l for l in result.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, result.stdout[-2000:]
    line = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
        assert key in line, key
    assert line["value"] > 0 and line["e2e"]["value"] > 0 and line["gpu_launches"] >= 20
    assert line["roofline"]["bound"] == "hbm" and line["roofline"]["achieved"] > 0
    assert "also" not in line            # secondary workloads only ride along with the default headline workload


_RANK_WORKER = r'''
import os, sys, types, json
root, rank, world, port = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
sys.path.insert(0, root)
import torch
import torch.distributed as dist
from metalbm_b200 import capi
if os.environ.get("MLBM_EMULATED") == "1":   # tests/conftest.py: the library compiled for the host
    capi._library = capi.load_library(os.environ["MLBM_EMULATED_LIBRARY"])
    torch.cuda.mem_get_info = lambda *a: (10 ** 11, 10 ** 11)
import bench
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
def reduce(op):
    def f(value):
        t = torch.tensor([float(value)], dtype=torch.float64)
        dist.all_reduce(t, op=op)
        return float(t.item())
    return f
for name, work in bench.WORKLOADS.items():
    work["shape"] = (8, 6, 5) if work["shape"][2] > 1 else (16, 10, 1)
    if work["store_every"]:
        work["store_every"] = 2
args = types.SimpleNamespace(overlap="On", halo="peer", variant=0)
maximum, minimum = reduce(dist.ReduceOp.MAX), (lambda v: -reduce(dist.ReduceOp.MAX)(-v))
results = []
for entry in bench.ALSO_MULTI:
    name, dtype, eps, mode, steps = entry
    results.append(bench.measure_also((name, dtype, eps, mode, 4), args, rank, world, rank, dist.barrier, maximum, minimum, 6500.0))
if rank == 0:
    print("RESULTS " + json.dumps(results))
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
'''


@pytest.mark.parametrize("world", [2, 4])
def test_secondary_workloads_of_the_bench_on_several_ranks(tmp_path, world):
    """bench.measure_also for every multi-GPU secondary workload (toy grids), one process per rank: a fresh context per entry --
    NCCL communicator, peer mappings, spectral plans made and torn down again and again -- which is what the driver's 1 -> 8 GPU
    scaling run does."""
    import json
    import os
    import socket
    import subprocess
    import sys
    from pathlib import Path

    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    root = Path(__file__).resolve().parent.parent
    script = tmp_path / "worker.py"
    script.write_text(_RANK_WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = [subprocess.Popen([sys.executable, str(script), str(root), str(r), str(world), str(port)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True, env=dict(os.environ, OMP_NUM_THREADS="1")) for r in range(world)]
    outputs = []
    for r, p in enumerate(procs):
        try:
            out, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            out, _ = p.communicate()
        assert p.returncode == 0 and f"ok {r}" in out, out[-3000:]
        outputs.append(out)
    results = json.loads([l for l in outputs[0].splitlines() if l.startswith("RESULTS ")][0][len("RESULTS "):])
    import bench
    assert [r["name"] for r in results] == [e[0] for e in bench.ALSO_MULTI]
    for r in results:
        if "skipped" in r:      # x extent of the toy grid not divisible by the rank count
            continue
        assert r["value"] > 0 and r["halo"] == "peer" and abs(r["mass_per_node"] - 1.0) < 1e-3 and r["energy"] > 0, r


_HEADLINE_WORKER = r'''
import json, os, sys, types
root, rank, world, port = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
sys.path.insert(0, root)
os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
import torch
import torch.distributed as dist
from metalbm_b200 import capi
if os.environ.get("MLBM_EMULATED") == "1":   # tests/conftest.py: the library compiled for the host; no CUDA in this process
    capi._library = capi.load_library(os.environ["MLBM_EMULATED_LIBRARY"])
    torch.cuda.mem_get_info = lambda *a: (10 ** 11, 10 ** 11)
    torch.cuda.set_device = lambda *a: None
    torch.cuda.synchronize = lambda *a: None
    real_init, real_tensor, real_empty = dist.init_process_group, torch.tensor, torch.empty
    dist.init_process_group = lambda backend=None, device_id=None, **k: real_init("gloo", **k)
    torch.tensor = lambda *a, **k: real_tensor(*a, **{key: v for key, v in k.items() if key != "device"})
    torch.empty = lambda *a, **k: real_empty(*a, **{key: v for key, v in k.items() if key not in ("pin_memory", "device")})
import bench
for name, work in bench.WORKLOADS.items():
    work["shape"] = (8, 6, 5) if work["shape"][2] > 1 else (16, 10, 1)
    if work["store_every"]:
        work["store_every"] = 2
bench.ALSO_MULTI = [(n, d, e, m, 3) for n, d, e, m, _ in bench.ALSO_MULTI[:2]]
bench.ClockSampler.start = lambda self: None
args = types.SimpleNamespace(gpus=world, steps=4, warmup=3, impl="ours", edge=bench.EDGE, variant=0, workload=None, dtype="f64",
                             overlap="On", halo="peer", eps=None, store_every=None, no_e2e=False, no_cpu_baseline=True,
                             also="auto", also_timeout=600, shape=None, stored_mode=0)
assert bench.run_ours(args) == 0
print("ok", rank)
'''


@pytest.mark.parametrize("world", [2, 4])
def test_the_multi_gpu_headline_of_the_bench_on_several_ranks(tmp_path, world):
    """bench.run_ours as the driver launches it on N > 1 GPUs, on a toy box: the headline is BASELINE configs[4] strong-scaled
    (stored steps with whole fields + the spectral enstrophy on the analysis stream), the end-to-end leg reads the all-reduced
    observables back every step, the secondary workloads follow; ONE line from rank 0."""
    import json
    import os
    import socket
    import subprocess
    import sys
    from pathlib import Path

    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    root = Path(__file__).resolve().parent.parent
    script = tmp_path / "worker.py"
    script.write_text(_HEADLINE_WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = [subprocess.Popen([sys.executable, str(script), str(root), str(r), str(world), str(port)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True, env=dict(os.environ, OMP_NUM_THREADS="1")) for r in range(world)]
    outputs = []
    for r, p in enumerate(procs):
        try:
            out, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            out, _ = p.communicate()
        assert p.returncode == 0 and f"ok {r}" in out, out[-3000:]
        outputs.append(out)
    lines = [l for l in outputs[0].splitlines() if l.startswith("{")]
    assert len(lines) == 1 and not any(l.startswith("{") for out in outputs[1:] for l in out.splitlines())
    line = json.loads(lines[0])
    assert line["metric"] == "MLUPS (D3Q19, FP64)" and line["scaling"] == "strong" and line["n_gpus"] == world and line["value"] > 0
    assert line["config"]["name"] == "d3q19_bgk_1024" and line["config"]["store_every"] == 2
    assert line["config"]["stored_mode"].startswith("whole fields") and line["config"]["stored_steps_in_timed_region"] == 2
    assert line["config"]["stored_step_ms"] >= 0 and line["config"]["stored_step_with_analysis_ms"] >= 0
    assert line["config"]["halo"].startswith("direct peer stores") and line["roofline"]["traffic"] is None
    assert line["e2e"]["value"] > 0 and "error" not in line["e2e"] and line["e2e"]["last_energy"] > 0
    assert [e["name"] for e in line["also"]] == ["d3q19_bgk_256", "d3q19_bgk_512"] and len(line["also_summary"]) == 2
    assert all(e.get("value", 0) > 0 for e in line["also"]), line["also"]


_STAGED_WORKER = r'''
import os, sys
root, out = sys.argv[1], sys.argv[2]
sys.path.insert(0, root); sys.path.insert(0, root + "/tests")
import numpy as np
from metalbm_b200 import capi
if os.environ.get("MLBM_EMULATED") == "1":
    capi._library = capi.load_library(os.environ["MLBM_EMULATED_LIBRARY"])
from metalbm_b200.algorithm import Algorithm
from metalbm_b200.capi import make_config
from oracle import oracle as O
results = {}
for key, lattice, shape, dtype in (("a", "D2Q9", (9, 11, 1), "F64"), ("b", "D3Q19", (5, 4, 7), "F64"), ("c", "D3Q27", (3, 4, 130), "F32")):
    cfg = make_config(lattice=lattice, shape=shape, forcing_scheme="Guo", force="Kolmogorov", tau=0.6, dtype=dtype,
                      amplitude=(1e-4, 1e-4, 1e-4), wavelength=(4.0, 4.0, 4.0))
    f0 = O.synthetic_populations(cfg, eps=1e-2)
    with Algorithm(cfg) as algorithm:
        algorithm.distribution.array[...] = 7.0           # padding included
        algorithm.distribution.set_interior(f0.astype(algorithm.domain.dtype))
        algorithm.unpack()
        algorithm.distribution.array[...] = -1.0
        algorithm.pack()
        results[key + "_roundtrip"] = algorithm.distribution.array.copy()
        for iteration in (1, 2):
            algorithm.iterate(iteration)
        algorithm.pack()
        results[key + "_stepped"] = algorithm.distribution.get_interior()
np.savez(out, **results)
print("ok")
'''


def test_staged_pack_and_unpack_equal_the_pitched_copies(tmp_path):
    """MLBM_STAGED_COPY=1 (experiment, off by default): every population crosses PCIe as one contiguous padded block and the
    padding is stripped / added by a kernel.  Same interior values as the pitched cudaMemcpy2DAsync route, bit for bit; the
    padding comes back as zeros where the pitched route leaves the host's values alone."""
    import os
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    script = tmp_path / "worker.py"
    script.write_text(_STAGED_WORKER)
    outputs = {}
    for mode in ("0", "1"):
        out = tmp_path / f"staged{mode}.npz"
        result = subprocess.run([sys.executable, str(script), str(root), str(out)], capture_output=True, text=True, timeout=600,
                                env=dict(os.environ, MLBM_STAGED_COPY=mode, OMP_NUM_THREADS="1"))
        assert result.returncode == 0 and "ok" in result.stdout, result.stdout[-2000:] + result.stderr[-2000:]
        outputs[mode] = np.load(out)
    for key in ("a", "b", "c"):
        pitched, staged = outputs["0"][key + "_roundtrip"], outputs["1"][key + "_roundtrip"]
        assert np.array_equal(outputs["0"][key + "_stepped"], outputs["1"][key + "_stepped"])
        interior = pitched != -1.0                              # the pitched route writes the interior only
        assert interior.any() and not interior.all()
        assert np.array_equal(staged[interior], pitched[interior])
        assert np.all(staged[~interior] == 0.0)


def test_alpha_statistics_follow_the_alpha_field():
    """mlbm_alpha_statistics: share of nodes off the alpha = 2 shortcut, min and max alpha, against the oracle's alpha field."""
    from helpers import run_oracle
    cfg = make_config(lattice="D2Q9", shape=(16, 140, 1), collision="ELBM", forcing_scheme="Guo", force="Kolmogorov", tau=0.55,
                      amplitude=(1e-4, 1e-4, 1e-4), wavelength=(8.0, 8.0, 8.0))
    f0 = O.synthetic_populations(cfg, eps=1e-5, amplitude=0.0, ripple=0.0)
    rng = np.random.default_rng(5)
    for _ in range(60):
        f0[:, rng.integers(0, 16), rng.integers(0, 140), 0] *= 1.0 + 0.05 * rng.standard_normal(9)
    with Algorithm(cfg) as algorithm:
        algorithm.distribution.set_interior(f0)
        algorithm.unpack()
        algorithm.iterate(1)
        fraction, low, high = algorithm.alpha_statistics()
    ref = run_oracle(cfg, f0, 1)
    assert 0.0 < fraction < 0.5
    assert abs(fraction - (ref.alpha != 2.0).mean()) <= 2.0 / ref.alpha.size
    assert abs(low - ref.alpha.min()) <= 1e-9 and abs(high - ref.alpha.max()) <= 1e-9
    with Algorithm(make_config(lattice="D2Q9", shape=(8, 8, 1))) as bgk:
        assert bgk.alpha_statistics() == (0.0, 2.0, 2.0)


def test_newton_statistics_count_the_solves_of_a_step():
    """mlbm_newton_statistics: nodes that took the Newton solve and the evaluations of (F, F') they needed in the steps between
    the two calls, against the oracle's branch and iteration bookkeeping (solveAlpha, Collision.h:328-349; EntropicStep.h:111-140).
    A node within rounding of a branch threshold may fall on the other side: a budget of 1 % of the nodes."""
    from helpers import run_oracle
    for lattice, shape in (("D2Q9", (12, 140, 1)), ("D3Q27", (6, 5, 9))):
        cfg = make_config(lattice=lattice, shape=shape, collision="ELBM", forcing_scheme="Guo", force="Kolmogorov", tau=0.55,
                          amplitude=(1e-4, 1e-4, 1e-4), wavelength=(8.0, 8.0, 8.0))
        f0 = O.synthetic_populations(cfg, eps=2e-2)
        with Algorithm(cfg) as algorithm:
            algorithm.distribution.set_interior(f0)
            algorithm.unpack()
            assert algorithm.newton_statistics(start=True) == (0, 0)
            algorithm.iterate(1)
            solved, evaluations = algorithm.newton_statistics()
            assert algorithm.newton_statistics() == (0, 0)      # counting stopped with the read
        ref = run_oracle(cfg, f0, 1)
        newton = ref.branch >= 2
        budget = max(2, int(0.01 * newton.size))
        assert abs(solved - int(newton.sum())) <= budget
        expected = int(ref.iterations[newton].sum())
        assert abs(evaluations - expected) <= max(4 * budget, 0.05 * expected)   # a convergence test within rounding of 1e-8 costs one more
        assert solved > 0.5 * newton.size and evaluations >= solved
    with Algorithm(make_config(lattice="D2Q9", shape=(8, 8, 1))) as bgk:
        assert bgk.newton_statistics(start=True) == (0, 0) and bgk.newton_statistics() == (0, 0)


@pytest.mark.parametrize("lattice,shape", [("D2Q9", (5, 7, 1)), ("D3Q19", (4, 3, 5)), ("D2Q13", (6, 5, 1))])
def test_halo_space_download_is_the_periodic_padding_of_the_distribution(lattice, shape):
    """mlbm_download_halo_distribution: Distribution::getHaloDataPrevious() as a host array in the reference's halo space hSD
    (Domain.h:173-283) -- local extents + 2 dimH per used dimension, every halo cell the periodic image (one rank)."""
    import ctypes
    from metalbm_b200.capi import check
    cfg = make_config(lattice=lattice, shape=shape, tau=0.6)
    f0 = O.synthetic_populations(cfg, eps=1e-2)
    dim, q, celerity, _ = O.lattice(lattice)
    halo = int(np.abs(celerity).max())
    with Algorithm(cfg) as algorithm:
        algorithm.distribution.set_interior(f0)
        algorithm.unpack()
        extents = [n + 2 * halo if d < dim else 1 for d, n in enumerate(shape)]
        out = np.full([q] + extents, np.nan)
        check(algorithm._lib.mlbm_download_halo_distribution(algorithm._ctx, out.ctypes.data, out.size))
        pad = [(0, 0)] + [(halo, halo) if d < dim else (0, 0) for d in range(3)]
        assert np.array_equal(out, np.pad(f0, pad, mode="wrap"))
        with pytest.raises(RuntimeError, match="needs dimQ"):
            check(algorithm._lib.mlbm_download_halo_distribution(algorithm._ctx, out.ctypes.data, out.size - 1))
