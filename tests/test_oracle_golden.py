"""The CPU oracle (oracle/lbm_oracle.c) against the golden vectors the reference itself produced
(tests/golden/make_golden.py): bit-for-bit on populations, alpha, density and force."""
import numpy as np
import pytest

from golden_util import golden_names, load_golden
from helpers import energy_removal_direct, native_shell_config, shell_force_direct
from metalbm_b200.capi import make_config
from oracle import oracle as O


@pytest.mark.parametrize("name", golden_names())
def test_oracle_reproduces_reference_bit_for_bit(name, oracle_lib):
    meta, cfg, data = load_golden(name)
    state = O.OracleState(cfg, data["f0"])
    if meta["force"] == "Field":
        # the force array the reference itself filled (ConstantShell / Turbulent2D, Force.h:296-623), read back through the
        # generic array read (Force.h:39-48)
        state.force[...] = data["force"]
    observed = []
    for _ in range(meta["steps"]):
        state.step(True)
        observed.append(state.observables())
    if meta.get("native_spectral"):
        # EnergyRemoval / Turbulent2D: the force array is remade every step from the stored fields with two transforms
        # (Force.h:452-550); numpy's FFT and the reference's (the oracle's DFT stub of FFTW) round differently
        assert np.abs(state.force - data["force"]).max() <= 4e-15 * np.abs(data["force"]).max()
        assert np.abs(state.f - data["f"]).max() <= 1e-14 * np.abs(data["f"]).max()
        assert np.abs(state.alpha - data["alpha"]).max() <= 1e-10
        assert np.abs(state.density - data["density"]).max() <= 1e-14
    else:
        assert np.array_equal(state.f, data["f"]), "populations differ from the reference"
        assert np.array_equal(state.alpha, data["alpha"]), "alpha differs from the reference"
        assert np.array_equal(state.density, data["density"])
        assert np.array_equal(state.force, data["force"])
    golden_observables = data["observables"]
    if meta["ranks"] == 1:
        # the reference's stored velocity went through its in-place FFT round trip (Routine.h:129-132): 1e-16 noise
        assert np.abs(state.velocity - data["velocity"]).max() <= 1e-15
    for row, mine in zip(golden_observables, observed):
        assert abs(mine[0] - row[1]) <= 1e-13 * abs(row[1])          # total energy (Analysis.h:53-61)
        if meta["ranks"] == 1:                                       # spectral enstrophy needs the single-rank FFT stub
            assert abs(mine[1] - row[2]) <= 1e-11 * abs(row[2])      # total enstrophy (Analysis.h:85-93)


def test_every_alpha_branch_is_covered_by_the_golden_set(oracle_lib):
    seen = set()
    for name in golden_names():
        meta, cfg, data = load_golden(name)
        if meta["collision"] == "BGK":
            continue
        state = O.OracleState(cfg, data["f0"])
        if meta["force"] == "Field":
            state.force[...] = data["force"]
        state.step(True)
        seen |= set(np.unique(state.branch).tolist())
    assert {0, 1, 2, 3} <= seen


SHELL_GOLDEN = ["d2q9_bgk_guo_constantshell", "d2q9_elbm_edm_constantshell", "d2q9_bgk_shanchen_turbulent2d"]


@pytest.mark.parametrize("name", SHELL_GOLDEN)
def test_constant_shell_restatement_against_the_reference_arrays(name, oracle_lib):
    """oracle.constant_shell_force (Force.h:296-420 + Transformer.h:300-384 restated) against the force array the reference
    itself made, and the oracle run from the NATIVE ConstantShell configuration against the reference's populations."""
    meta, _, data = load_golden(name)
    cfg = native_shell_config(meta)
    scale = np.abs(data["force"]).max()
    assert np.abs(O.constant_shell_force(cfg) - data["force"]).max() <= 2e-15 * scale
    assert np.abs(shell_force_direct(cfg) - data["force"]).max() <= 4e-15 * scale       # the device kernel's formula
    state = O.OracleState(cfg, data["f0"])
    for _ in range(meta["steps"]):
        state.step(True)
    assert np.abs(state.f - data["f"]).max() <= 1e-14 * np.abs(data["f"]).max()
    assert np.abs(state.alpha - data["alpha"]).max() <= 1e-10


@pytest.mark.parametrize("shape", [(8, 6, 1), (7, 5, 1), (8, 5, 1), (9, 8, 1)])
@pytest.mark.parametrize("shell", [(1, 2), (0, 4), (2, 6), (3, 3)])
def test_constant_shell_direct_sum_equals_the_transform(shape, shell):
    """The device synthesises the shell force as a sum over the shell's modes instead of a c2r transform: equal for even and
    odd extents and for shells that reach the Nyquist wave numbers (where the transform drops non-Hermitian parts)."""
    cfg = make_config("D2Q9", shape, force="ConstantShell", amplitude=(1e-4, 0.0, 0.0), k_min=shell[0], k_max=shell[1])
    transform = O.constant_shell_force(cfg)
    assert np.abs(transform).max() > 0
    assert np.abs(shell_force_direct(cfg) - transform).max() <= 4e-15 * np.abs(transform).max()


@pytest.mark.parametrize("shape", [(8, 6, 1), (7, 5, 1), (9, 8, 1)])
@pytest.mark.parametrize("shell", [(1, 2), (0, 4), (2, 6)])
def test_energy_removal_mode_sums_equal_the_transforms(shape, shell, oracle_lib):
    """The device evaluates EnergyRemoval as a projection onto the shell's modes and a synthesis (csrc/shell_force.cu) instead of
    the reference's r2c / c2r pair: equal for even and odd extents and for shells that reach the Nyquist wave numbers."""
    cfg = make_config("D2Q9", shape, force="EnergyRemoval", amplitude=(2e-3, 3e-3, 0.0), k_min=shell[0], k_max=shell[1])
    rng = np.random.default_rng(7)
    density = 1.0 + 0.1 * rng.standard_normal(shape)
    velocity = 0.05 * rng.standard_normal((2,) + shape)
    transform = O.energy_removal_force(cfg, density, velocity, cfg.force_amplitude, shell[0], shell[1])
    direct = energy_removal_direct(cfg, density, velocity, cfg.force_amplitude, shell[0], shell[1])
    assert np.abs(transform).max() > 0
    assert np.abs(direct - transform).max() <= 1e-13 * np.abs(transform).max()


@pytest.mark.parametrize("name", [n for n in golden_names()])
def test_power_spectra_restatement_against_the_reference(name, oracle_lib):
    """oracle.power_spectra (SpectralAnalysisList / PowerSpectra restated, maxWaveNumber quirk included) against the spectra
    the reference computed from the very fields stored in the golden file."""
    meta, cfg, data = load_golden(name)
    if "spectra" not in data.files:
        pytest.skip("recorded on several ranks: the reference's FFT stub is single-rank")
    spectra = data["spectra"]
    assert spectra.shape == (O.max_wave_number(cfg), 2)
    mine = O.power_spectra(cfg, data["velocity"], data["force"])
    assert np.abs(mine[:, 0] - spectra[:, 0]).max() <= 1e-13 * np.abs(spectra[:, 0]).max()
    assert np.abs(mine[:, 1] - spectra[:, 1]).max() <= 1e-13 * max(np.abs(spectra[:, 1]).max(), 1e-300)


def test_max_wave_number_follows_the_reference_quirk():
    """arrayMax is Max(first, arrayMIN of the rest) (Helpers.h:37-39)."""
    assert O.max_wave_number(make_config("D3Q19", (8, 6, 10))) == 4
    assert O.max_wave_number(make_config("D3Q19", (6, 8, 10))) == 4
    assert O.max_wave_number(make_config("D2Q9", (12, 16, 1))) == 6
