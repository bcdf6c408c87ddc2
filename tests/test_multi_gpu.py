"""x-slab decomposition over several GPUs: the decomposition must be invisible.

One process per GPU (the reference's one-MPI-rank-per-GPU model, CUDAInitializer.h:23-26); the 128-byte NCCL id is
shipped over a gloo process group, the halo planes travel GPU to GPU inside mlbm_step.  Each rank advances its
slab; rank files are gathered and compared with (a) the single-rank CPU oracle on the same global field, within
BASELINE.json's tolerances, and (b) the single-GPU CUDA result, which must be BIT-IDENTICAL (same kernel, same
per-node arithmetic, only the origin of the halo values differs)."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from helpers import relative_error, run_cuda, run_oracle
from metalbm_b200.capi import make_config
from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent

WORKER = r'''
import json, sys
import numpy as np
import torch
import torch.distributed as dist
root, rank, world, port, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
sys.path.insert(0, root)
import os
from metalbm_b200 import capi
if os.environ.get("MLBM_EMULATED") == "1":   # tests/conftest.py: the library compiled for the host, NCCL over shared memory
    capi._library = capi.load_library(os.environ["MLBM_EMULATED_LIBRARY"])
from metalbm_b200.algorithm import Algorithm, Communication, slab_of
from metalbm_b200.capi import make_config

case = json.load(open(out + "/case.json"))
f0 = np.load(out + "/f0.npy")
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
cfg = make_config(rank=rank, nranks=world, device=rank, **case["config"])
algorithm = Algorithm(cfg, communication=Communication(rank, world), peer_halos=case["peer"])
assert algorithm.peer_halos == case["peer"]
domain = algorithm.domain
algorithm.distribution.set_interior(slab_of(f0, rank, world).astype(domain.dtype))
algorithm.unpack()
steps = case["steps"]
if case["mode"] == "sync":
    for iteration in range(1, steps + 1):
        algorithm.isStored = iteration == steps
        algorithm.iterate(iteration)
elif case["mode"] == "stored":   # every step stored: the spectral forces follow the stored fields
    for iteration in range(1, steps + 1):
        algorithm.isStored = True
        algorithm.iterate(iteration)
else:  # the asynchronous run loop with a stored step at the end
    algorithm.run(1, steps - 1)
    algorithm.isStored = True
    algorithm.iterate(steps)
observables = algorithm.observables()
if case.get("checkpoint"):
    algorithm.write_checkpoint(out + "/checkpoint.mlbm", steps)   # every rank its hyperslab of the same file
algorithm.pack()
fields = algorithm.fieldList
np.savez(out + f"/rank{rank}.npz", f=algorithm.distribution.get_interior(), density=domain.interior(fields.density)[0],
         velocity=domain.interior(fields.velocity), alpha=domain.interior(fields.alpha)[0], observables=observables)
algorithm.close()
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
'''


def _device_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _run_ranks(tmp_path, world, config, f0, steps, mode, peer=False, checkpoint=False):
    import json
    (tmp_path / "case.json").write_text(json.dumps({"config": config, "steps": steps, "mode": mode, "peer": peer, "checkpoint": checkpoint}))
    np.save(tmp_path / "f0.npy", f0)
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    procs = [subprocess.Popen([sys.executable, str(script), str(ROOT), str(r), str(world), str(port), str(tmp_path)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env) for r in range(world)]
    for r, p in enumerate(procs):
        try:
            out, _ = p.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            out, _ = p.communicate()
        assert p.returncode == 0 and f"ok {r}" in out, out[-3000:]
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    return {
        "f": np.concatenate([p["f"] for p in parts], axis=1),
        "density": np.concatenate([p["density"] for p in parts], axis=0),
        "velocity": np.concatenate([p["velocity"] for p in parts], axis=1),
        "alpha": np.concatenate([p["alpha"] for p in parts], axis=0),
        "observables": [p["observables"] for p in parts],
    }


CASES = [
    # lattice, shape, collision, scheme, force, overlap, steps, mode[, halo]
    ("D3Q19", (16, 6, 10), "BGK", "Guo", "Kolmogorov", "On", 5, "sync", "peer"),
    ("D3Q19", (16, 6, 10), "BGK", "None", "None", "On", 12, "async", "peer"),
    ("D2Q9", (24, 20, 1), "BGK", "Guo", "Kolmogorov", "On", 9, "async", "peer"),
    ("D3Q27", (8, 6, 4), "ELBM", "Guo", "Kolmogorov", "On", 2, "sync", "peer"),
    ("D2Q9", (8, 140, 1), "ELBM", "ShanChen", "Kolmogorov", "On", 3, "async", "peer"),
    ("D3Q19", (16, 6, 10), "BGK", "Guo", "Kolmogorov", "On", 5, "sync"),
    ("D3Q19", (16, 6, 10), "BGK", "Guo", "Kolmogorov", "Off", 5, "sync"),
    ("D3Q19", (16, 6, 10), "BGK", "None", "None", "On", 6, "async"),
    ("D2Q9", (24, 20, 1), "BGK", "Guo", "Kolmogorov", "On", 5, "async"),
    ("D3Q27", (8, 6, 4), "ELBM", "Guo", "Kolmogorov", "On", 2, "sync"),
    ("D2Q9", (8, 12, 1), "ELBM", "ExactDifferenceMethod", "Kolmogorov", "Off", 3, "async"),
]


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(map(str, (c[0], c[2], c[3], c[5], c[7]) + c[8:])))
def test_slabs_reproduce_the_single_rank_result(tmp_path, world, case):
    if _device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    lattice, shape, collision, scheme, force, overlap, steps, mode = case[:8]
    peer = len(case) > 8 and case[8] == "peer"   # boundary kernel stores into the neighbours' halo planes (CUDA IPC)
    if shape[0] % world or shape[0] // world < 1:
        pytest.skip("slab too thin")
    config = dict(lattice=lattice, shape=list(shape), collision=collision, forcing_scheme=scheme, force=force, tau=0.55,
                  amplitude=[1e-4, 2e-4, 3e-4], wavelength=[8.0, 4.0, 16.0], overlap=overlap)
    single = make_config(**config)
    f0 = O.synthetic_populations(single, eps=1e-2)
    got = _run_ranks(tmp_path, world, config, f0, steps, mode, peer)

    # (b) bit-identical to the single-GPU CUDA path.  The force profile is evaluated at LOCAL x (Collision.h:86), which
    # only Sinusoidal forces along x would notice; the cases here are x-independent.
    one = run_cuda(single, f0, steps)
    assert np.array_equal(got["f"], one["f"])
    assert np.array_equal(got["alpha"], one["alpha"])
    assert np.array_equal(got["density"], one["density"])

    # (a) the oracle, BASELINE tolerances
    ref = run_oracle(single, f0, steps)
    if collision == "BGK":
        assert relative_error(got["f"], ref.f) <= 1e-12 * steps
    obs = ref.observables()
    for rank_observables in got["observables"]:   # every rank holds the reduced values
        assert abs(rank_observables[0] - obs[0]) <= 1e-9 * abs(obs[0])
        assert abs(rank_observables[1] - obs[1]) <= 1e-9 * abs(obs[1])   # spectral enstrophy, distributed transform
        assert abs(rank_observables[3] - obs[3]) <= 1e-12 * abs(obs[3])
        assert abs(rank_observables[2] - obs[2]) <= 1e-12 * abs(obs[2])
        assert np.array_equal(rank_observables, got["observables"][0])


@pytest.mark.parametrize("world", [2, 4])
def test_checkpoint_written_on_slabs_restarts_on_one_rank(tmp_path, world):
    """DistributionWriter / DistributionReader (Writer.h:400-445, Reader.h:119-157): `world` ranks write their hyperslabs of the
    dimQ padded-global-box data sets into ONE file; a single-rank context reads the whole box back, bit for bit."""
    if _device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from metalbm_b200.algorithm import Algorithm
    config = dict(lattice="D3Q19", shape=[16, 6, 10], collision="BGK", forcing_scheme="Guo", force="Kolmogorov", tau=0.55,
                  amplitude=[1e-4, 2e-4, 3e-4], wavelength=[8.0, 4.0, 16.0], overlap="On")
    single = make_config(**config)
    f0 = O.synthetic_populations(single, eps=1e-2)
    got = _run_ranks(tmp_path, world, config, f0, 3, "sync", peer=True, checkpoint=True)
    with Algorithm(single) as algorithm:
        assert algorithm.read_checkpoint(tmp_path / "checkpoint.mlbm") == 3
        algorithm.pack()
        assert np.array_equal(algorithm.distribution.get_interior(), got["f"])
