"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares,
refuses to run without a GPU (no CPU fallback), and its host logic (domain arithmetic, halo plan) is right.
The two-rank halo plan is exercised over torch.distributed's gloo backend (world_size 2)."""
import ctypes
import os
import re
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from metalbm_b200 import capi
from metalbm_b200.algorithm import Communication, Domain, slab_of

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol(cuda_lib):
    header = (ROOT / "include" / "metalbm_b200.h").read_text()
    declared = set(re.findall(r"\b(mlbm_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(capi.PROTOTYPES), "ctypes prototypes and header out of sync"
    for name in declared:
        assert hasattr(cuda_lib, name), f"{name} not exported"
    assert cuda_lib.mlbm_abi_version() == capi.ABI_VERSION


def test_struct_layout_matches_header(cuda_lib):
    assert ctypes.sizeof(capi.MlbmConfig) == 8 * 4 + 3 * 4 + 4 * 4 + 4 + 8 + 24 + 24 + 4 * 4 + 24  # includes 4 bytes of padding before tau
    assert ctypes.sizeof(capi.MlbmHaloMessage) == 32


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not _no_gpu(), reason="a GPU is present")
def test_create_fails_loudly_without_a_gpu(cuda_lib):
    cfg = capi.make_config("D3Q19", (8, 8, 8))
    ctx = ctypes.c_void_p()
    status = cuda_lib.mlbm_create(ctypes.byref(cfg), ctypes.byref(ctx))
    assert status == -2 and not ctx.value
    assert b"no CPU fallback" in cuda_lib.mlbm_last_error()


def test_create_rejects_bad_configurations(cuda_lib):
    ctx = ctypes.c_void_p()
    bad = [capi.make_config("D3Q19", (8, 8, 8), equilibrium="Exact"),      # Exact exists for D2Q9 / D3Q27 only
           capi.make_config("D3Q19", (9, 8, 8), nranks=2),                 # numProcs must divide globalLengthX
           capi.make_config("D2Q9", (8, 8, 1), tau=0.5),
           capi.make_config("D2Q17", (4, 8, 1), nranks=2),                 # slabs of 2 planes under a halo of 3
           capi.make_config("D3Q33", (8, 8, 8), equilibrium="Exact"),
           capi.make_config("D3Q19", (8, 8, 8), force="ConstantShell"),    # the shell force is rebuilt for 2-D lattices only
           capi.make_config("D2Q9", (8, 8, 1), force="ConstantShell", k_min=3, k_max=2)]
    for cfg in bad:
        assert cuda_lib.mlbm_create(ctypes.byref(cfg), ctypes.byref(ctx)) == -1
    cfg = capi.make_config("D2Q9", (8, 8, 1))
    cfg.abi_version = 99
    assert cuda_lib.mlbm_create(ctypes.byref(cfg), ctypes.byref(ctx)) == -1


def test_domain_padding_follows_the_reference():
    # lSD::pLength pads the LAST used dimension to 2 (n/2 + 1) (Domain.h:53-57, MathVector.h:330-344)
    d3 = Domain(capi.make_config("D3Q19", (8, 6, 5), nranks=2, rank=1))
    assert d3.local_length == (4, 6, 5) and d3.padded_length == (4, 6, 6) and d3.number_elements == 144 and d3.offset_x == 4
    d2 = Domain(capi.make_config("D2Q9", (8, 6, 1)))
    assert d2.local_length == (8, 6, 1) and d2.padded_length == (8, 8, 1)


def test_halo_plan_single_rank_is_empty(cuda_lib):
    assert capi.halo_plan(capi.make_config("D3Q19", (8, 8, 8))) == []


@pytest.mark.parametrize("lattice,shape,face", [("D2Q9", (8, 6, 1), 3), ("D3Q19", (8, 6, 4), 5), ("D3Q27", (8, 6, 4), 9)])
def test_halo_plan_messages(cuda_lib, lattice, shape, face):
    cfg = capi.make_config(lattice, shape, nranks=4, rank=1)
    plan = capi.halo_plan(cfg)
    assert len(plan) == 4 * face
    sends = [m for m in plan if m.is_send]
    assert sorted(m.population for m in sends) == list(range(1, 2 * face + 1))
    for m in sends:
        assert m.peer == (0 if m.population <= face else 2)   # c_x < 0 travels left, c_x > 0 right
    plane = shape[1] * shape[2]
    assert all(m.count == plane for m in plan)


@pytest.mark.parametrize("lattice,shape,face,halo", [("D2Q13", (8, 6, 1), 4, 2), ("D2Q21", (12, 6, 1), 7, 3), ("D3Q33", (8, 6, 4), 10, 2)])
def test_halo_plan_of_the_multi_speed_lattices(cuda_lib, lattice, shape, face, halo):
    """dimH planes per side travel together (Communication.h:145-150, sizeStripeX): the last H interior planes of the c_x > 0
    populations to the right neighbour's planes 0..H-1, the first H interior planes of the c_x < 0 ones to the left neighbour's
    planes LX+H..LX+2H-1; interior planes are H..LX+H-1 of a population."""
    cfg = capi.make_config(lattice, shape, nranks=2, rank=1)
    plan = capi.halo_plan(cfg)
    lx, plane = shape[0] // 2, shape[1] * shape[2]
    stride = capi.launch_plan(cfg, 0, lx).stride
    assert stride >= plane * (lx + 2 * halo) and len(plan) == 4 * face and all(m.count == halo * plane for m in plan)
    for m in plan:
        first_plane = (m.offset - m.population * stride) // plane
        right_going = m.population > face
        if m.is_send:
            assert first_plane == (lx if right_going else halo)
        else:
            assert first_plane == (0 if right_going else lx + halo)


WORKER = r'''
import sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from metalbm_b200 import capi
from metalbm_b200.algorithm import Communication, slab_of
from oracle import oracle as O

rank, world, port = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
lattice, shape = "D3Q19", (8, 6, 4)
cfg = capi.make_config(lattice, shape, nranks=world, rank=rank)
dim, q, c, w = O.lattice(lattice)

# the 128-byte id hand-shake of Algorithm.__init__ over gloo
communication = Communication(rank, world)
payload = communication.broadcast_bytes(bytes(range(128)) if rank == 0 else None, 128)
assert payload == bytes(range(128))
assert communication.rank_left == (rank - 1) % world and communication.rank_right == (rank + 1) % world

# a global field every rank can rebuild; the slab in the device layout [Q][LX + 2][plane]
rng = np.random.default_rng(5)
full = rng.standard_normal((q,) + shape)
lx, plane = shape[0] // world, shape[1] * shape[2]
perPopulation = (lx + 2) * plane
stride = (perPopulation + 31) // 32 * 32
buffer = np.zeros(q * stride)
for iq in range(q):
    buffer[iq * stride + plane: iq * stride + (lx + 1) * plane] = slab_of(full[iq], rank, world).ravel()

# execute the plan with point-to-point messages (what exchangeHalos does with ncclSend / ncclRecv)
plan = capi.halo_plan(cfg)
requests, receives = [], []
for m in plan:
    if m.is_send:
        t = torch.from_numpy(buffer[m.offset:m.offset + m.count].copy())
        requests.append(dist.isend(t, dst=m.peer, tag=m.population))
    else:
        t = torch.zeros(m.count, dtype=torch.float64)
        requests.append(dist.irecv(t, src=m.peer, tag=m.population))
        receives.append((m, t))
for r in requests:
    r.wait()
for m, t in receives:
    buffer[m.offset:m.offset + m.count] = t.numpy()

# every population that pulls across a slab face now finds its periodic upstream neighbour in the halo plane
x0 = rank * lx
for iq in range(q):
    cx = c[iq, 0]
    if cx == 1:
        got = buffer[iq * stride: iq * stride + plane]
        assert np.array_equal(got, full[iq, (x0 - 1) % shape[0]].ravel()), f"left halo of population {iq}"
    if cx == -1:
        got = buffer[iq * stride + (lx + 1) * plane: iq * stride + (lx + 2) * plane]
        assert np.array_equal(got, full[iq, (x0 + lx) % shape[0]].ravel()), f"right halo of population {iq}"
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
'''


@pytest.mark.parametrize("world", [2, 4])
def test_halo_plan_over_gloo(cuda_lib, oracle_lib, tmp_path, world):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    procs = [subprocess.Popen([sys.executable, str(script), str(ROOT), str(r), str(world), str(port)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env) for r in range(world)]
    outputs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            p.kill()
            out, _ = p.communicate()
        outputs.append(out)
    for r, (p, out) in enumerate(zip(procs, outputs)):
        assert p.returncode == 0 and f"ok {r}" in out, out[-2000:]


# ---------------------------------------------------------------------------------------------------------------
# The scalar kernel parameters and the grid of a launch, as mlbm_step assembles them (device-free mirror).  A launch
# whose scalars are silently wrong (beta = 0, no periodic wrap, stored flag dropped) runs at full speed and produces
# plausible-looking numbers: this is the CPU-side guard.
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("lattice,shape,collision,scheme,force,tau,nranks", [
    ("D3Q19", (256, 256, 256), "BGK", "None", "None", 0.55, 1),
    ("D3Q19", (1024, 1024, 1024), "BGK", "Guo", "Kolmogorov", 0.55, 8),
    ("D3Q27", (512, 512, 512), "ELBM", "Guo", "Kolmogorov", 0.50000032, 1),
    ("D2Q9", (8192, 8192, 1), "ELBM", "ShanChen", "Kolmogorov", 0.7, 8),
    ("D2Q9", (24, 130, 1), "ForcedNR_ELBM", "ExactDifferenceMethod", "Constant", 0.9, 2),
    ("D3Q27", (16, 12, 10), "BGK", "Guo", "Field", 0.6, 2),            # array-type forces read the force field
    ("D2Q17", (64, 48, 1), "ELBM", "Guo", "Kolmogorov", 0.6, 1),       # multi-speed: inv_cs2 = 2/3 in Guo's prefactor
    ("D3Q33", (16, 12, 10), "BGK", "Guo", "Kolmogorov", 0.6, 1),
    ("D2Q9", (64, 48, 1), "ELBM", "Guo", "ConstantShell", 0.6, 4),
])
def test_launch_plan_scalars_follow_the_configuration(cuda_lib, lattice, shape, collision, scheme, force, tau, nranks):
    cfg = capi.make_config(lattice=lattice, shape=shape, collision=collision, forcing_scheme=scheme, force=force, tau=tau,
                           nranks=nranks, rank=nranks - 1, overlap="On")
    dim, q = capi.LATTICE_DQ[capi.Lattice(cfg.lattice)]
    lx = shape[0] // nranks
    nm, nr = (shape[1], shape[2]) if dim == 3 else (1, shape[1])
    for is_stored in (0, 1, 2):
        plan = capi.launch_plan(cfg, 0, lx, is_stored)
        assert plan.beta == 1.0 / (2.0 * tau)                                  # Collision.h:122
        inv_cs2 = capi.LATTICE_INV_CS2.get(capi.Lattice(cfg.lattice), 3.0)
        assert plan.guo_factor == (1.0 - 1.0 / (2.0 * tau)) * inv_cs2          # ForcingScheme.h:115
        assert plan.wrap_x == (1 if nranks == 1 else 0)
        assert plan.is_stored == is_stored
        assert plan.hydro_shift == (0 if scheme == "None" else 1)              # ForcingScheme.h:26-33 / :50-57
        assert plan.has_force == (0 if force == "None" else (2 if force in ("Field", "ConstantShell") else 1))
        assert list(plan.local_length) == [lx, nm, nr]
        assert plan.plane == nm * nr and plan.stride >= plan.plane * (lx + 2) and plan.stride % 32 == 0
        assert plan.block == 128 and plan.grid[0] == -(-nr // 128) and plan.grid[1] == nm
        assert plan.x0 == 0 and plan.plane_step == 1 and plan.plane_count == lx
        # every plane is covered exactly once
        assert plan.grid[2] == -(-lx // plan.planes_per_block)
        entropic = collision != "BGK"
        assert (plan.shared_bytes > 2 * q * 128 * 8) == entropic
        assert plan.shared_bytes <= 227 * 1024
        if not entropic:
            assert plan.planes_per_block == 1
        else:
            assert 1 <= plan.planes_per_block <= 16
            # the grid keeps at least ~20 waves of 4 blocks on 148 SMs whenever blocks walk several planes
            assert plan.planes_per_block == 1 or plan.grid[0] * plan.grid[1] * plan.grid[2] >= 148 * 4 * 20 // 2


def test_launch_plan_of_the_two_boundary_planes(cuda_lib):
    """Overlapped multi-GPU steps compute planes 0 and LX - 1 in one launch (plane_step = LX - 1), one plane per block."""
    cfg = capi.make_config(lattice="D3Q27", shape=(512, 512, 512), collision="ELBM", forcing_scheme="Guo", force="Kolmogorov",
                           tau=0.55, nranks=8, rank=3, overlap="On")
    lx = 64
    plan = capi.launch_plan(cfg, 0, 2, 1, lx - 1)
    assert (plan.x0, plan.plane_step, plan.plane_count, plan.planes_per_block, plan.grid[2]) == (0, lx - 1, 2, 1, 2)
    assert plan.wrap_x == 0 and plan.is_stored == 1 and plan.beta == 1.0 / 1.1
    bulk = capi.launch_plan(cfg, 1, lx - 1, 0)
    assert bulk.x0 == 1 and bulk.plane_count == lx - 2 and bulk.grid[2] * bulk.planes_per_block >= lx - 2


def test_launch_plan_rejects_bad_ranges(cuda_lib):
    cfg = capi.make_config(lattice="D2Q9", shape=(16, 12, 1), tau=0.7)
    plan = capi.MlbmLaunchPlan()
    for x0, x1, step in ((0, 0, 1), (-1, 4, 1), (0, 17, 1), (0, 2, 16), (3, 2, 1)):
        assert cuda_lib.mlbm_launch_plan_for(ctypes.byref(cfg), x0, x1, 0, step, ctypes.byref(plan)) == -1


def test_header_is_plain_c_and_links_from_c(tmp_path, cuda_lib):
    """include/metalbm_b200.h compiles as C99 with -pedantic and the library links from a C program (what a cgo / JNI / Fortran
    binding relies on); without a device the program reports the loud mlbm_create failure."""
    binary = tmp_path / "capi_minimal"
    command = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", str(ROOT / "include"), str(ROOT / "examples" / "capi_minimal.c"),
               "-L", str(ROOT / "metalbm_b200"), "-lmetalbm_b200", f"-Wl,-rpath,{ROOT / 'metalbm_b200'}", "-o", str(binary)]
    result = subprocess.run(command, capture_output=True, text=True)
    assert result.returncode == 0, result.stderr[-3000:]
    run = subprocess.run([str(binary)], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, run.stdout + run.stderr
    assert ("no CPU fallback" in run.stdout) if _no_gpu() else run.stdout.startswith("energy ")
