"""GPU parity of the array-type forces: the generic force-array read (Force "Field",
Force.h:39-48) and the spectral forces rebuilt for 2-D lattices (ConstantShell, EnergyRemoval, Turbulent2D, Force.h:296-616),
against golden vectors of the reference run with those forces and against the oracle.  Kept in a file of its own that sorts
after the established parity suites: these are the newest device code paths, and a run that stops at the first failure must
not lose the rest to them."""
import numpy as np
import pytest

from golden_util import golden_names
from helpers import check_entropic, force_field, native_shell_config, relative_error, run_cuda, run_oracle
from metalbm_b200.capi import make_config
from oracle import oracle as O
from test_cpp_shim import SPECTRAL_SHIM_CASES, check_template_api, compile_example
from test_golden_gpu import check_cuda_against_golden
from test_multi_gpu import _device_count, _run_ranks
from test_parity_gpu import ENERGY_TOLERANCE, POPULATION_TOLERANCE, _config, _flow

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_names(spectral=True, wide=False))
def test_cuda_reproduces_reference_outputs_with_spectral_forces(name):
    check_cuda_against_golden(name)


FIELD_FORCE_CASES = [
    # lattice, shape, equilibrium, scheme, tau, collision, eps
    ("D2Q9", (33, 130, 1), "TruncationMa3", "Guo", 0.7, "BGK", 1e-2),
    ("D3Q19", (12, 10, 9), "TruncationMa3", "ExactDifferenceMethod", 0.6, "BGK", 1e-2),
    ("D3Q27", (8, 6, 4), "Exact", "ShanChen", 0.55, "BGK", 1e-2),
    ("D2Q9", (16, 140, 1), "TruncationMa3", "Guo", 0.55, "ELBM", 2e-2),
    ("D3Q19", (8, 6, 4), "TruncationMa3", "Guo", 0.55, "ForcedNR_ELBM_Forcing", 2e-2),
]


@pytest.mark.parametrize("case", FIELD_FORCE_CASES, ids=lambda c: "-".join(map(str, c[:1] + c[2:4] + c[5:6])))
def test_force_read_from_the_force_field(case):
    """Force "Field" (mlbm_set_force_field): the generic array read Force<Generic>::setForce (Force.h:39-48) that the
    reference's spectral forces run through; pinned to the reference by tests/golden/*constantshell*.npz."""
    lattice, shape, equilibrium, scheme, tau, collision, eps = case
    cfg = _config(lattice, shape, equilibrium, scheme, "Field", tau, collision)
    f0 = O.synthetic_populations(cfg, eps=eps, **_flow(eps))
    field = force_field(cfg)
    for steps in (1, 3):
        got = run_cuda(cfg, f0, steps, force=field)
        ref = run_oracle(cfg, f0, steps, force=field)
        if collision == "BGK":
            assert relative_error(got["f"], ref.f) <= POPULATION_TOLERANCE
            assert np.abs(got["velocity"] - ref.velocity).max() <= 1e-13
        else:
            check_entropic(got, ref, cfg, steps, mismatch_budget=1e-3)
        assert np.array_equal(got["force"], field)      # storeFields writes the values it read back (Algorithm.h:186-190)


def test_field_force_equals_the_analytic_force_bit_for_bit():
    """The Kolmogorov profile handed over as an array: the same doubles through the other route."""
    analytic = _config("D3Q19", (12, 10, 130), "TruncationMa3", "Guo", "Kolmogorov", 0.55)
    f0 = O.synthetic_populations(analytic, eps=1e-2)
    want = run_cuda(analytic, f0, 3)
    array = _config("D3Q19", (12, 10, 130), "TruncationMa3", "Guo", "Field", 0.55)
    got = run_cuda(array, f0, 3, force=want["force"])
    assert np.array_equal(got["f"], want["f"]) and np.array_equal(got["velocity"], want["velocity"])


def test_force_field_needs_a_field_force_context():
    from metalbm_b200.algorithm import Algorithm
    from metalbm_b200.capi import MlbmError
    with Algorithm(_config("D2Q9", (8, 8, 1), "TruncationMa3", "Guo", "Kolmogorov", 0.7)) as algorithm:
        with pytest.raises(MlbmError):
            algorithm.set_force()


@pytest.mark.parametrize("name", ["d2q9_bgk_guo_constantshell", "d2q9_elbm_edm_constantshell", "d2q9_bgk_shanchen_turbulent2d"])
def test_native_constant_shell_against_the_reference(name):
    """Force "ConstantShell" synthesised on the device at mlbm_create (csrc/shell_force.cu: injectionKernel) against golden
    vectors of the reference run with its own ConstantShell / Turbulent2D force: the force array and the populations."""
    from golden_util import load_golden
    meta, _, data = load_golden(name)
    cfg = native_shell_config(meta)
    got = run_cuda(cfg, data["f0"], meta["steps"])
    assert np.abs(got["force"] - data["force"]).max() <= 2e-14 * np.abs(data["force"]).max()   # device sinpi: a few ulp per mode
    if meta["collision"] == "BGK":
        assert relative_error(got["f"], data["f"]) <= POPULATION_TOLERANCE
    else:
        ref = run_oracle(cfg, data["f0"], meta["steps"])
        check_entropic(got, ref, cfg, meta["steps"], mismatch_budget=5e-3)
    energy = data["observables"][-1][1]
    assert abs(got["observables"][0] - energy) <= ENERGY_TOLERANCE * abs(energy)


@pytest.mark.parametrize("shape,shell,dtype", [((33, 20, 1), (1, 2), "F64"), ((16, 15, 1), (0, 9), "F64"), ((24, 130, 1), (2, 3), "F32")])
def test_native_constant_shell_odd_sizes_and_nyquist_shells(shape, shell, dtype):
    cfg = make_config(lattice="D2Q9", shape=shape, collision="BGK", forcing_scheme="Guo", force="ConstantShell", tau=0.7,
                      amplitude=(2e-3, 0.0, 0.0), k_min=shell[0], k_max=shell[1], dtype=dtype)
    f0 = O.synthetic_populations(cfg, eps=1e-2)
    got = run_cuda(cfg, f0, 2)
    ref = run_oracle(cfg, f0, 2)
    scale = np.abs(ref.force).max()
    assert scale > 0 and np.abs(got["force"] - ref.force).max() <= (1e-13 if dtype == "F64" else 1e-7) * scale   # up to ~150 modes, a few ulp of the device sinpi each
    assert relative_error(got["f"], ref.f) <= (POPULATION_TOLERANCE if dtype == "F64" else 1e-5)


SPECTRAL_FORCE_CASES = [
    # force, shape, collision, scheme, dtype, shell keywords
    ("EnergyRemoval", (33, 20, 1), "BGK", "Guo", "F64", dict(amplitude=(2e-3, 3e-3, 0.0), k_min=1, k_max=3)),
    ("EnergyRemoval", (16, 15, 1), "ELBM", "ExactDifferenceMethod", "F64", dict(amplitude=(4e-3, 1e-3, 0.0), k_min=0, k_max=9)),
    ("Turbulent2D", (24, 130, 1), "BGK", "Guo", "F64", dict(amplitude=(2e-4, 0.0, 0.0), k_min=1, k_max=2,
                                                             removal_amplitude=(5e-3, 2e-3, 0.0), removal_k_min=2, removal_k_max=4)),
    ("Turbulent2D", (24, 20, 1), "BGK", "ShanChen", "F32", dict(amplitude=(2e-4, 0.0, 0.0), k_min=1, k_max=2,
                                                                removal_amplitude=(5e-3, 2e-3, 0.0), removal_k_min=1, removal_k_max=3)),
]


@pytest.mark.parametrize("case", SPECTRAL_FORCE_CASES, ids=lambda c: "-".join(map(str, (c[0], "x".join(map(str, c[1])), c[2], c[3], c[4]))))
def test_time_dependent_spectral_forces(case):
    """EnergyRemoval / Turbulent2D (Force.h:423-616) on the device: the force follows the fields of the last STORED step.
    Every step is stored here (the array changes every step); pinned to the reference by tests/golden/*energyremoval*,
    *turbulent2d_removal*."""
    force, shape, collision, scheme, dtype, shell = case
    cfg = make_config(lattice="D2Q9", shape=shape, collision=collision, forcing_scheme=scheme, force=force, tau=0.6, dtype=dtype, **shell)
    f0 = O.synthetic_populations(cfg, eps=1e-2)
    from metalbm_b200.algorithm import Algorithm
    with Algorithm(cfg) as algorithm:
        domain = algorithm.domain
        algorithm.distribution.set_interior(f0.astype(domain.dtype))
        algorithm.unpack()
        state = O.OracleState(cfg, f0)
        for iteration in range(1, 5):
            algorithm.isStored = True
            algorithm.iterate(iteration)
            state.step(True)
            got_force = domain.interior(algorithm.fieldList.force).astype(np.float64)
            scale = max(np.abs(state.force).max(), 1e-30)
            assert np.abs(got_force - state.force).max() <= (1e-11 if dtype == "F64" else 1e-6) * scale, f"step {iteration}"
        algorithm.pack()
        got = algorithm.distribution.get_interior().astype(np.float64)
    assert np.abs(state.force).max() > 1e-6
    if collision == "BGK":
        assert relative_error(got, state.f) <= (4 * POPULATION_TOLERANCE if dtype == "F64" else 1e-5)
    else:
        alpha = domain.interior(algorithm.fieldList.alpha)[0].astype(np.float64)
        check_entropic({"f": got, "alpha": alpha}, state, cfg, 4, mismatch_budget=5e-3)


def test_spectral_force_follows_the_stored_fields_only():
    """Between stored steps fieldList does not change, so neither does the reference's EnergyRemoval array (Force.h:552-558
    recomputes it from the same fields): steps 1-3 unstored, step 4 stored, steps 5-6 unstored."""
    cfg = make_config(lattice="D2Q9", shape=(20, 18, 1), collision="BGK", forcing_scheme="Guo", force="EnergyRemoval", tau=0.7,
                      amplitude=(3e-3, 3e-3, 0.0), k_min=1, k_max=3)
    f0 = O.synthetic_populations(cfg, eps=1e-2)
    from metalbm_b200.algorithm import Algorithm
    with Algorithm(cfg) as algorithm:
        algorithm.distribution.set_interior(f0)
        algorithm.unpack()
        state = O.OracleState(cfg, f0)
        for iteration in range(1, 7):
            algorithm.isStored = iteration == 4
            algorithm.iterate(iteration)
            state.step(iteration == 4)
        algorithm.pack()
        got = algorithm.distribution.get_interior()
    assert np.abs(state.force).max() > 1e-6          # the force switched on after the stored step
    assert relative_error(got, state.f) <= 6 * POPULATION_TOLERANCE


POWER_SPECTRA_GOLDEN = ["d2q9_bgk_guo_kolmogorov", "d2q9_bgk_none", "d3q15_bgk_edm", "d3q19_bgk_none", "d3q19_bgk_guo_kolmogorov",
                        "d3q27_elbm_guo", "d2q9_bgk_guo_constantshell"]


@pytest.mark.parametrize("name", POWER_SPECTRA_GOLDEN)
def test_power_spectra_against_the_reference(name):
    """mlbm_power_spectra (energy spectrum of the stored velocity, forcing spectrum of the force array; distributed cuFFT
    transform + binning kernel) against the spectra the reference's SpectralAnalysisList computed (golden vectors)."""
    from golden_util import load_golden
    from metalbm_b200.algorithm import Algorithm
    meta, cfg, data = load_golden(name)
    with Algorithm(cfg) as algorithm:
        algorithm.distribution.set_interior(data["f0"])
        algorithm.unpack()
        if meta["force"] == "Field":
            algorithm.domain.interior(algorithm.fieldList.force)[...] = data["force"]
            algorithm.set_force()
        for iteration in range(1, meta["steps"] + 1):
            algorithm.isStored = iteration == meta["steps"]
            algorithm.iterate(iteration)
        got = algorithm.power_spectra()
    spectra = data["spectra"]
    assert got.shape == spectra.shape
    assert np.abs(got[:, 0] - spectra[:, 0]).max() <= 1e-9 * np.abs(spectra[:, 0]).max()
    assert np.abs(got[:, 1] - spectra[:, 1]).max() <= 1e-9 * max(np.abs(spectra[:, 1]).max(), 1e-300)


@pytest.mark.parametrize("lattice,shape", [("D2Q9", (33, 20, 1)), ("D2Q9", (16, 131, 1)), ("D3Q19", (9, 8, 7)), ("D3Q27", (6, 8, 10))])
def test_power_spectra_odd_sizes_against_the_oracle(lattice, shape):
    from metalbm_b200.algorithm import Algorithm
    cfg = _config(lattice, shape, "TruncationMa3", "Guo", "Sinusoidal", 0.6)
    f0 = O.synthetic_populations(cfg, eps=1e-2)
    with Algorithm(cfg) as algorithm:
        algorithm.distribution.set_interior(f0)
        algorithm.unpack()
        algorithm.isStored = True
        algorithm.iterate(1)
        got = algorithm.power_spectra()
        again = algorithm.power_spectra()               # the fields are left untouched
    ref = run_oracle(cfg, f0, 1)
    want = O.power_spectra(cfg, ref.velocity, ref.force)
    assert got.shape == want.shape and np.array_equal(got.shape, again.shape)
    assert np.abs(got - want).max(axis=0)[0] <= 1e-10 * np.abs(want[:, 0]).max()
    assert np.abs(got - want).max(axis=0)[1] <= 1e-10 * np.abs(want[:, 1]).max()
    assert np.abs(got - again).max() <= 1e-12 * np.abs(got).max()


@pytest.mark.parametrize("world", [1, 2])
@pytest.mark.parametrize("case", SPECTRAL_SHIM_CASES, ids=lambda c: "-".join(map(str, c[:1] + c[2:6])))
def test_template_api_with_spectral_forces(tmp_path, cuda_lib, world, case):
    check_template_api(tmp_path, world, case)


@pytest.mark.parametrize("peer", [True, False], ids=["peer", "nccl"])
@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("force", ["ConstantShell", "Turbulent2D"])
def test_spectral_forces_on_slabs(tmp_path, world, force, peer):
    """ConstantShell / Turbulent2D on x-slabs: the synthesis uses global coordinates, the projection of the stored momentum
    onto the shell's modes is all-reduced over the ranks (csrc/shell_force.cu).  ConstantShell must be bit-identical to the
    single-GPU run; with the removal part the summation order of the projection differs, so the oracle's tolerance applies."""
    if _device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    shape, steps = (24, 20, 1), 4
    config = dict(lattice="D2Q9", shape=list(shape), collision="BGK", forcing_scheme="Guo", force=force, tau=0.6,
                  amplitude=[2e-4, 0.0, 0.0], overlap="On", k_min=1, k_max=2, removal_amplitude=[5e-3, 2e-3, 0.0],
                  removal_k_min=2, removal_k_max=4)
    single = make_config(**config)
    f0 = O.synthetic_populations(single, eps=1e-2)
    got = _run_ranks(tmp_path, world, config, f0, steps, "stored", peer)
    if force == "ConstantShell":
        one = run_cuda(single, f0, steps, store_every_step=True)
        assert np.array_equal(got["f"], one["f"])
    ref = run_oracle(single, f0, steps)
    assert relative_error(got["f"], ref.f) <= 1e-12 * steps
    assert abs(got["observables"][0][0] - ref.observables()[0]) <= 1e-9 * abs(ref.observables()[0])


def test_reference_style_routine_writes_the_spectra_table(tmp_path, cuda_lib):
    """src/main.cu with Architecture::GPU and spectralAnalysisStep = 50: Routine::compute appends the energy / forcing spectra
    to ../output/<prefix>/spectra_<startIteration>.dat (SpectralAnalysisList, AnalysisList.h:99-202; SpectralAnalysisWriter,
    Writer.h:193-251: "iteration wavenumber energy_spectra forcing_spectra"), against the oracle's restatement."""
    import subprocess
    shape = (32, 24, 1)
    binary = compile_example(tmp_path, "main_gpu.cpp", "main_gpu", "D2Q9", shape, scheme="Guo", force="Kolmogorov", tau=0.55,
                             steps=100, spectral_step=50)
    run = tmp_path / "run"
    run.mkdir()
    result = subprocess.run([str(binary)], capture_output=True, text=True, cwd=run, timeout=300)
    assert result.returncode == 0, result.stdout[-2000:] + result.stderr[-2000:]
    lines = (tmp_path / "output" / "test" / "spectra_0.dat").read_text().splitlines()
    assert lines[0] == "iteration wavenumber energy_spectra forcing_spectra"
    table = np.array([[float(v) for v in line.split()] for line in lines[1:]])
    cfg = make_config(lattice="D2Q9", shape=shape, collision="BGK", forcing_scheme="Guo", force="Kolmogorov", tau=0.55,
                      amplitude=(1e-4, 2e-4, 3e-4), wavelength=(8.0, 4.0, 16.0))
    bins = O.max_wave_number(cfg)
    assert table.shape == (2 * bins, 4) and all(line.endswith(" ") for line in lines[1:])
    assert table[:, 0].tolist() == [50.0] * bins + [100.0] * bins and table[:bins, 1].tolist() == list(map(float, range(bins)))
    density = np.ones(shape)
    density[int((shape[0] - 1) * 0.4), int((shape[1] - 1) * 0.3), 0] = 3.0     # initDensity Peak (Initialize.h:30-46)
    state = O.OracleState(cfg, O.init_equilibrium(cfg, density, np.zeros((2,) + shape)))
    for iteration in range(1, 101):
        state.step(iteration % 50 == 0)
        if iteration % 50 == 0:
            want = O.power_spectra(cfg, state.velocity, state.force)
            got = table[(iteration // 50 - 1) * bins:(iteration // 50) * bins, 2:]
            assert np.abs(got - want).max(axis=0)[0] <= 1e-9 * np.abs(want[:, 0]).max()
            assert np.abs(got - want).max(axis=0)[1] <= 1e-9 * np.abs(want[:, 1]).max()
