"""The C++ drop-in layer (include/metaLBM_b200/metaLBM/*.h): reference spellings over the C-ABI.

CPU part: the headers compile with g++ -std=c++14 against an Input.in in the reference's format, the lattice / domain
descriptors equal the reference's tables (through the oracle, which restates Lattice.h), and the reference-style
main links against libmetalbm_b200.so.  GPU part: populations injected through Distribution, advanced with
Algorithm::iterate and read back with Algorithm::pack match the oracle within BASELINE.json's tolerances, on one rank
and on two ranks (processes) exchanging halos over NVLink."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from helpers import relative_error, run_oracle
from metalbm_b200.capi import make_config
from oracle import oracle as O

ROOT = Path(__file__).resolve().parent.parent
INCLUDE = ROOT / "include" / "metaLBM_b200"
LIBDIR = Path(os.environ.get("MLBM_SHIM_LIBDIR", ROOT / "metalbm_b200"))   # tests/conftest.py re-points it for emulated runs


def compile_example(tmp_path, source, name, lattice, shape, collision="BGK", equilibrium="TruncationMa3", scheme="Guo",
                    force="Kolmogorov", tau=0.55, nprocs=1, overlap="Off", link=True, input_file="Input_generic.in", steps=100,
                    compile_only=False, spectral_step=0, defines=()):
    output = tmp_path / name
    command = ["g++", "-std=c++14", "-O1", "-Wall", "-Wextra", "-Werror",
               f"-DNPROCS={nprocs}", "-DNTHREADS=1", f"-DGLOBAL_LENGTH_X={shape[0]}", f"-DGLOBAL_LENGTH_Y={shape[1]}",
               f"-DGLOBAL_LENGTH_Z={shape[2]}", '-DLBM_POSTFIX="test"', f"-DLBM_LATTICE={lattice}", f"-DLBM_COLLISION={collision}",
               f"-DLBM_EQUILIBRIUM={equilibrium}", f"-DLBM_SCHEME={scheme}", f"-DLBM_FORCE={force}", f"-DLBM_TAU={tau}",
               f"-DLBM_OVERLAP={overlap}", f"-DLBM_STEPS={steps}", f"-DLBM_SPECTRAL_STEP={spectral_step}", *defines,
               "-include", str(ROOT / "examples" / input_file), "-I", str(INCLUDE), str(ROOT / "examples" / source),
               "-o", str(output)]
    if compile_only:
        command.insert(1, "-c")
    elif link:
        command += ["-L", str(LIBDIR), "-lmetalbm_b200", f"-Wl,-rpath,{LIBDIR}"]
    result = subprocess.run(command, capture_output=True, text=True)
    assert result.returncode == 0, result.stderr[-4000:]
    return output


@pytest.mark.parametrize("lattice,shape", [("D2Q5", (8, 6, 1)), ("D2Q9", (8, 6, 1)), ("D3Q15", (8, 6, 4)),
                                           ("D3Q19", (8, 6, 4)), ("D3Q27", (8, 6, 5)), ("D2Q13", (8, 6, 1)), ("D2Q17", (8, 6, 1)),
                                           ("D2Q21", (8, 6, 1)), ("D3Q33", (8, 6, 5))])
def test_lattice_descriptor_equals_reference_tables(tmp_path, oracle_lib, lattice, shape):
    binary = compile_example(tmp_path, "lattice_dump.cpp", "lattice_dump", lattice, shape, link=False)
    lines = subprocess.run([str(binary)], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    dim, q, c, w = O.lattice(lattice)
    head = list(map(int, lines[0].split()))
    face = {"D2Q5": 1, "D2Q9": 3, "D3Q15": 5, "D3Q19": 5, "D3Q27": 9, "D2Q13": 4, "D2Q17": 7, "D2Q21": 7, "D3Q33": 10}[lattice]
    halo = int(np.abs(c).max())                      # dimH = the longest jump (Lattice.h:223, 300, 382, 716)
    assert head == [dim, q, halo, face]
    for iq in range(q):
        parts = lines[1 + iq].split()
        assert list(map(int, parts[:3])) == c[iq].tolist()
        assert float(parts[3]) == w[iq]
    faces = [list(map(int, part.split())) for part in lines[1 + q].split("|")]
    expect = [[i for i in range(q) if c[i, axis] * sign > 0] if axis < dim else [] for axis, sign in ((1, -1), (1, 1), (2, -1), (2, 1))]
    assert faces == expect
    # lSD::pLength pads the last used dimension to 2 (n / 2 + 1) (Domain.h:53-57); hSD::volume has a halo of 1 per used side
    px, py, pz, pvolume, hvolume = map(int, lines[2 + q].split())
    padded = list(shape)
    padded[dim - 1] = 2 * (shape[dim - 1] // 2 + 1)
    assert [px, py, pz] == padded and pvolume == int(np.prod(padded))
    assert hvolume == int(np.prod([n + 2 * halo if i < dim else n for i, n in enumerate(shape)]))
    assert int(lines[3 + q]) == 2 ** 32 - 1


def test_reference_style_main_compiles_and_links(tmp_path, cuda_lib):
    compile_example(tmp_path, "main_gpu.cpp", "main_gpu", "D2Q9", (64, 64, 1), input_file="Input_d2q9_kolmogorov.in")
    compile_example(tmp_path, "shim_check.cpp", "shim_check", "D3Q27", (8, 6, 4), collision="ELBM", equilibrium="Exact")


def test_unsupported_choices_fail_at_compile_time(tmp_path):
    """static_asserts of the template layer, not link errors: compiled with -c."""
    for force in ("ConstantShell", "EnergyRemoval", "Turbulent2D"):
        compile_example(tmp_path, "shim_check.cpp", "shell.o", "D2Q9", (24, 20, 1), force=force, compile_only=True)
    compile_example(tmp_path, "shim_check.cpp", "alpha.o", "D3Q19", (8, 6, 4), collision="Malaspinas_ELBM", compile_only=True)
    rejected = [dict(lattice="D3Q19", shape=(8, 6, 4), equilibrium="Exact"),          # Exact: D2Q9 / D3Q27 only (Equilibrium.h:36-126)
                dict(lattice="D3Q19", shape=(8, 6, 4), force="ConstantShell"),        # the shell force is rebuilt for 2-D lattices only
                dict(lattice="D3Q27", shape=(8, 6, 4), force="EnergyRemoval"),
                dict(lattice="D3Q15", shape=(8, 6, 4), force="Turbulent2D")]
    for choice in rejected:
        with pytest.raises(AssertionError, match="static assertion failed"):
            compile_example(tmp_path, "shim_check.cpp", "bad.o", choice.pop("lattice"), choice.pop("shape"), compile_only=True, **choice)


def _run_shim(tmp_path, binary, f0_slabs, steps, env_extra=None):
    procs = []
    world = len(f0_slabs)
    for rank, slab in enumerate(f0_slabs):
        slab.astype(np.float64).tofile(tmp_path / f"in{rank}.bin")
        env = dict(os.environ, MLBM_RANK=str(rank), MLBM_NRANKS=str(world), MLBM_RENDEZVOUS_DIR=str(tmp_path),
                   MLBM_SESSION=f"shim{os.getpid()}")
        env.update(env_extra or {})
        procs.append(subprocess.Popen([str(binary), str(tmp_path / f"in{rank}.bin"), str(steps), str(tmp_path / f"out{rank}.bin"),
                                       str(tmp_path / f"fields{rank}.bin"), str(tmp_path / f"moments{rank}.bin")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                      text=True, env=env))
    outputs = []
    for rank, p in enumerate(procs):
        try:
            out, _ = p.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            for other in procs:
                other.kill()
            out, _ = p.communicate()
        assert p.returncode == 0 and f"ok rank {rank}" in out, out[-3000:]
        outputs.append(out)
    return outputs


SHIM_CASES = [
    ("D3Q19", (16, 12, 10), "BGK", "TruncationMa3", "Guo", "Kolmogorov", 3),
    ("D2Q9", (24, 20, 1), "BGK", "TruncationMa3", "ExactDifferenceMethod", "Kolmogorov", 3),
    ("D3Q27", (8, 6, 4), "ELBM", "TruncationMa3", "Guo", "Kolmogorov", 1),
]
# the spectral force types run in tests/test_spectral_forces_gpu.py (same check, later in the order)
SPECTRAL_SHIM_CASES = [
    ("D2Q9", (24, 20, 1), "BGK", "TruncationMa3", "Guo", "ConstantShell", 3),   # forcekMin / forcekMax of examples/Input_generic.in
    ("D2Q9", (24, 20, 1), "BGK", "TruncationMa3", "Guo", "Turbulent2D", 3),     # + removalForce* of examples/Input_generic.in
]


@pytest.mark.gpu
@pytest.mark.parametrize("world", [1, 2])
@pytest.mark.parametrize("case", SHIM_CASES, ids=lambda c: "-".join(map(str, c[:1] + c[2:6])))
def test_template_api_reproduces_the_oracle(tmp_path, cuda_lib, world, case):
    check_template_api(tmp_path, world, case)


def check_template_api(tmp_path, world, case):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    lattice, shape, collision, equilibrium, scheme, force, steps = case
    binary = compile_example(tmp_path, "shim_check.cpp", "shim_check", lattice, shape, collision, equilibrium, scheme, force,
                             tau=0.55, nprocs=world, overlap="On" if world > 1 else "Off")
    cfg = make_config(lattice=lattice, shape=shape, collision=collision, equilibrium=equilibrium, forcing_scheme=scheme,
                      force=force, tau=0.55, amplitude=(1e-4, 2e-4, 3e-4), wavelength=(8.0, 4.0, 16.0), k_min=1, k_max=2,
                      removal_amplitude=(5e-3, 2e-3, 0.0), removal_k_min=2, removal_k_max=4)   # examples/Input_generic.in
    f0 = O.synthetic_populations(cfg, eps=1e-2)
    lx = shape[0] // world
    outputs = _run_shim(tmp_path, binary, [np.ascontiguousarray(f0[:, r * lx:(r + 1) * lx]) for r in range(world)], steps)
    ref = run_oracle(cfg, f0, steps, store_last_only=True)   # examples/shim_check.cpp stores the last step only
    dim, q = ref.dim, ref.q
    got = np.concatenate([np.fromfile(tmp_path / f"out{r}.bin").reshape((q, lx) + tuple(shape[1:])) for r in range(world)], axis=1)
    if collision == "BGK":
        assert relative_error(got, ref.f) <= 1e-12 * steps
    else:
        node_error = np.abs(got - ref.f).max(axis=0)
        alpha_ok = np.ones(shape, dtype=bool)
    volume = lx * shape[1] * shape[2]
    for r in range(world):
        raw = np.fromfile(tmp_path / f"fields{r}.bin")
        density = raw[:volume].reshape((lx,) + tuple(shape[1:]))
        alpha = raw[(1 + dim) * volume:(2 + dim) * volume].reshape((lx,) + tuple(shape[1:]))
        observables = raw[(2 + dim) * volume:]
        assert relative_error(density, ref.density[r * lx:(r + 1) * lx]) <= 1e-12 * steps
        if collision == "BGK":
            assert np.all(alpha == 2.0)
        else:
            tolerance = 1e-10 + 4.0 * ref.alpha_noise[r * lx:(r + 1) * lx]
            alpha_ok[r * lx:(r + 1) * lx] = np.abs(alpha - ref.alpha[r * lx:(r + 1) * lx]) <= tolerance
        obs = ref.observables()
        assert abs(observables[0] - obs[0]) <= 1e-9 * abs(obs[0])
        assert abs(observables[3] - obs[3]) <= 1e-12 * abs(obs[3])
        # Communication::reduce of the stored density == the reduced mass observable
        mass = float(outputs[r].split("mass")[1].split()[0])
        assert abs(mass - obs[3]) <= 1e-12 * abs(obs[3])
    if collision != "BGK":
        assert (~alpha_ok).mean() <= 1e-3
        assert node_error[alpha_ok].max() <= 1e-12 * np.abs(ref.f).max() + 2e-10 * ref.fneq_max.max()
    # the host-callable per-node surface (Moment<T>, Collision::calculateMoments / setForce / getHydrodynamicVelocity) over the
    # halo-space host copy: moments of the populations PULLED to every node of the state after the last step (Moment.h:14-47),
    # against the same sums over the device result itself, periodic images from np.roll
    _, _, celerity, _ = O.lattice(lattice)
    pulled = np.stack([np.roll(got[iq], tuple(int(c) for c in celerity[iq]), axis=(0, 1, 2)) for iq in range(q)])
    density = pulled.sum(axis=0)
    velocity = np.stack([(pulled * celerity[:, d, None, None, None]).sum(axis=0) for d in range(dim)]) / density
    force = ref.force            # Kolmogorov: a profile along y, the same at local and global coordinates
    for r in range(world):
        raw = np.fromfile(tmp_path / f"moments{r}.bin").reshape((1 + 2 * dim, lx) + tuple(shape[1:]))
        part = slice(r * lx, (r + 1) * lx)
        assert relative_error(raw[0], density[part]) <= 1e-14
        assert np.abs(raw[1:1 + dim] - velocity[:, part]).max() <= 1e-14
        hydro = velocity[:, part] + (0.5 / density[part]) * force[:, part] if scheme != "None" else velocity[:, part]
        assert np.abs(raw[1 + dim:] - hydro).max() <= 1e-14


def test_observables_file_format_equals_the_reference(tmp_path):
    """`../output/<prefix>/observables_<startIteration>.dat` as the drop-in ScalarAnalysisWriter writes it against the
    committed bytes of the reference's own writer (Writer.h:140-190) and, where /root/reference is present, against the
    reference class compiled on the spot."""
    golden = (ROOT / "tests" / "golden" / "observables_reference_format.dat").read_bytes()
    run = tmp_path / "ours" / "run"
    run.mkdir(parents=True)
    binary = compile_example(tmp_path, "observables_writer.cpp", "observables_writer", "D2Q9", (8, 6, 1), link=False)
    assert subprocess.run([str(binary)], cwd=run).returncode == 0          # creates ../output/test/ itself
    mine = (tmp_path / "ours" / "output" / "test" / "observables_7.dat").read_bytes()
    assert mine == golden
    from oracle import refbuild
    if not refbuild.reference_available():
        return
    build = tmp_path / "reference"
    (build / "run").mkdir(parents=True)
    (build / "output" / "test").mkdir(parents=True)                         # the reference does not create it
    (build / "Input.in").write_text(refbuild.input_in(refbuild.RefConfig(lattice="D2Q9", nx=8, ny=6, nz=1)).replace(
        'constexpr auto prefix = "oracle";', 'constexpr auto prefix = "test";'))
    command = ["g++", "-std=c++14", "-O1", "-w", "-DUSE_FFTW", "-DNPROCS=1", "-DNTHREADS=1", "-DGLOBAL_LENGTH_X=8",
               "-DGLOBAL_LENGTH_Y=6", "-DGLOBAL_LENGTH_Z=1", '-DLBM_POSTFIX="test"', "-DOBSERVABLES_WRITER_REFERENCE",
               f"-I{ROOT / 'oracle' / 'shim'}", f"-I{build}", f"-I{refbuild.REFERENCE_ROOT / 'include'}",
               str(ROOT / "examples" / "observables_writer.cpp"), "-o", str(build / "writer")]
    result = subprocess.run(command, capture_output=True, text=True)
    assert result.returncode == 0, result.stderr[-3000:]
    assert subprocess.run([str(build / "writer")], cwd=build / "run").returncode == 0
    assert (build / "output" / "test" / "observables_7.dat").read_bytes() == golden


@pytest.mark.gpu
def test_reference_style_routine_runs(tmp_path, cuda_lib):
    """src/main.cu with Architecture::GPU: Routine::compute from a density peak, observables against the oracle."""
    shape = (64, 48, 1)
    binary = compile_example(tmp_path, "main_gpu.cpp", "main_gpu", "D2Q9", shape, scheme="Guo", force="Kolmogorov", tau=0.55,
                             steps=100)
    run = tmp_path / "run"
    run.mkdir()
    result = subprocess.run([str(binary)], capture_output=True, text=True, cwd=run, timeout=300)
    assert result.returncode == 0, result.stdout[-2000:] + result.stderr[-2000:]
    # the reference's file (../output/<prefix>/observables_<startIteration>.dat, three columns) and the B200 extras next to it
    table = np.loadtxt(tmp_path / "output" / "test" / "observables_0.dat", skiprows=1)
    extras = np.loadtxt(tmp_path / "output" / "test" / "observables_b200_0.dat", skiprows=1)
    assert table.shape == (2, 3) and extras.shape == (2, 3) and table[:, 0].tolist() == [50, 100] == extras[:, 0].tolist()
    # the same run on the oracle: rho = 1 with a 3x peak at (0.4, 0.3)(L - 1) (Initialize.h:30-46), u = 0, f = feq
    cfg = make_config(lattice="D2Q9", shape=shape, collision="BGK", forcing_scheme="Guo", force="Kolmogorov", tau=0.55,
                      amplitude=(1e-4, 2e-4, 3e-4), wavelength=(8.0, 4.0, 16.0))
    density = np.ones(shape)
    density[int((shape[0] - 1) * 0.4), int((shape[1] - 1) * 0.3), 0] = 3.0
    state = O.OracleState(cfg, O.init_equilibrium(cfg, density, np.zeros((2,) + shape)))
    for iteration in range(1, 101):
        state.step(iteration % 50 == 0)
        if iteration % 50 == 0:
            obs = state.observables()
            row = table[iteration // 50 - 1]
            assert abs(row[1] - obs[0]) <= 1e-9 * abs(obs[0])
            assert abs(row[2] - obs[1]) <= 1e-9 * abs(obs[1])
            assert abs(extras[iteration // 50 - 1][2] - obs[3]) <= 1e-12 * abs(obs[3])


@pytest.mark.gpu
def test_routine_backs_up_and_restarts_from_the_checkpoint(tmp_path, cuda_lib):
    """Routine::compute with backUpStep (Routine.h:212-216: pack + DistributionWriter) and a second binary with
    startIteration != 0 (initDistribution reads the checkpoint, Initialize.h:119-124): the restarted run writes the SAME final
    checkpoint as the uninterrupted one, bit for bit (BGK; the data sets are the reference's padded global box)."""
    import json
    shape = (24, 20, 1)
    common = dict(lattice="D2Q9", shape=shape, scheme="Guo", force="Kolmogorov", tau=0.6)
    whole = compile_example(tmp_path, "main_gpu.cpp", "whole", steps=8, defines=("-DLBM_BACKUP_STEP=4",), **common)
    resumed = compile_example(tmp_path, "main_gpu.cpp", "resumed", steps=8, defines=("-DLBM_BACKUP_STEP=4", "-DLBM_START=4"), **common)
    run = tmp_path / "run"
    run.mkdir()
    result = subprocess.run([str(whole)], capture_output=True, text=True, cwd=run, timeout=300)
    assert result.returncode == 0, result.stdout[-2000:] + result.stderr[-2000:]
    folder = tmp_path / "output" / "test"
    first = (folder / "distribution-8.mlbm").read_bytes()
    header = json.loads((folder / "distribution-4.mlbm").read_bytes()[:4096].decode())
    assert header["iteration"] == 4 and header["dimQ"] == 9 and header["padded_global_length"] == [24, 22, 1]
    (folder / "distribution-8.mlbm").unlink()
    result = subprocess.run([str(resumed)], capture_output=True, text=True, cwd=run, timeout=300)
    assert result.returncode == 0, result.stdout[-2000:] + result.stderr[-2000:]
    assert (folder / "distribution-8.mlbm").read_bytes() == first
    data = np.frombuffer(first[4096:], dtype=np.float64).reshape(9, 24, 22, 1)[:, :, :20]
    assert abs(data.sum() / (24 * 20) - 1.0) < 1e-2 and np.isfinite(data).all()      # a density peak of 3 on a background of 1
