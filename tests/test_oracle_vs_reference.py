"""Live check of the oracle against the reference compiled in this container (skipped where /root/reference and
the prebuilt oracle/_ref binary are both absent, e.g. on the GPU box)."""
import numpy as np
import pytest

from metalbm_b200.capi import make_config
from oracle import oracle as O
from oracle import refbuild

CASES = [
    # lattice, shape, collision, equilibrium, scheme, force, tau, eps, ranks
    ("D2Q9", (16, 12, 1), "BGK", "TruncationMa3", "Guo", "Kolmogorov", 0.7, 1e-2, 1),
    ("D3Q19", (8, 6, 4), "BGK", "TruncationMa3", "None", "None", 0.55, 1e-3, 1),
    ("D3Q19", (8, 6, 4), "BGK", "TruncationMa3", "Guo", "Kolmogorov", 0.55, 1e-3, 4),
    ("D3Q27", (8, 6, 4), "ELBM", "TruncationMa3", "Guo", "Kolmogorov", 0.55, 2e-2, 1),
    ("D2Q9", (16, 12, 1), "ELBM", "Exact", "ExactDifferenceMethod", "Kolmogorov", 0.51, 2e-2, 2),
    # alpha models whose overrides are dead code in the snapshot (== ELBM) and the one that is not
    ("D2Q9", (12, 10, 1), "Approached_ELBM", "TruncationMa3", "Guo", "Kolmogorov", 0.51, 2e-2, 1),
    ("D2Q9", (12, 10, 1), "Malaspinas_ELBM", "TruncationMa3", "Guo", "Kolmogorov", 0.51, 2e-2, 1),
    ("D2Q9", (12, 10, 1), "Essentially1_ELBM", "TruncationMa3", "Guo", "Kolmogorov", 0.51, 2e-2, 1),
    ("D2Q9", (12, 10, 1), "Essentially2_ELBM", "TruncationMa3", "Guo", "Kolmogorov", 0.51, 2e-2, 1),
    ("D2Q9", (12, 10, 1), "ForcedBNR_ELBM", "TruncationMa3", "Guo", "Kolmogorov", 0.51, 2e-2, 1),
    ("D2Q9", (12, 10, 1), "ForcedNR_ELBM_Forcing", "TruncationMa3", "Guo", "Kolmogorov", 0.51, 2e-2, 1),
    ("D3Q27", (6, 6, 4), "ForcedNR_ELBM_Forcing", "TruncationMa3", "ExactDifferenceMethod", "Kolmogorov", 0.55, 2e-2, 1),
    ("D3Q19", (6, 4, 4), "ForcedNR_ELBM_Forcing", "TruncationMa3", "Guo", "Kolmogorov", 0.55, 3e-1, 1),
    # multi-speed lattices (halos of 2-3 nodes, their own sound speeds)
    ("D2Q13", (12, 10, 1), "BGK", "TruncationMa3", "Guo", "Kolmogorov", 0.7, 1e-2, 1),
    ("D2Q17", (12, 10, 1), "BGK", "TruncationMa3", "ExactDifferenceMethod", "Kolmogorov", 0.7, 1e-2, 1),
    ("D2Q21", (12, 10, 1), "BGK", "TruncationMa3", "ShanChen", "Kolmogorov", 0.7, 1e-2, 1),
    ("D3Q33", (6, 5, 4), "BGK", "TruncationMa3", "Guo", "Kolmogorov", 0.6, 1e-2, 1),
    ("D2Q13", (12, 10, 1), "ELBM", "TruncationMa3", "Guo", "Kolmogorov", 0.55, 2e-2, 1),
    ("D2Q21", (12, 10, 1), "ELBM", "TruncationMa3", "Guo", "Kolmogorov", 0.55, 2e-2, 1),
    ("D3Q33", (6, 5, 4), "ELBM", "TruncationMa3", "ExactDifferenceMethod", "Kolmogorov", 0.55, 2e-2, 1),
    ("D2Q17", (12, 10, 1), "ForcedNR_ELBM_Forcing", "TruncationMa3", "Guo", "Kolmogorov", 0.55, 2e-2, 1),
]


def test_dead_alpha_overrides_equal_elbm_in_the_reference(oracle_lib):
    """Approached_, Malaspinas_, Essentially1/2_, ForcedNR_ and ForcedBNR_ELBM only override the private non-virtual
    calculateAlpha that Collision<ELBM>::calculateRelaxationTime never calls (Collision.h:239): the compiled reference
    produces bit-identical populations and alpha for all of them, which is why they share the ELBM kernel."""
    base = dict(lattice="D2Q9", nx=12, ny=10, nz=1, equilibrium="TruncationMa3", forcing_scheme="Guo", force="Kolmogorov", tau=0.51)
    _reference_or_skip(refbuild.RefConfig(collision="ELBM", **base))
    cfg = make_config(lattice="D2Q9", shape=(12, 10, 1), collision="ELBM", forcing_scheme="Guo", force="Kolmogorov", tau=0.51,
                      amplitude=(1e-5, 1e-5, 1e-5))
    f0 = O.synthetic_populations(cfg, eps=2e-2)
    expected = refbuild.run_ref(refbuild.RefConfig(collision="ELBM", **base), f0, 2)
    for collision in ("Approached_ELBM", "Malaspinas_ELBM", "Essentially1_ELBM", "Essentially2_ELBM", "ForcedNR_ELBM", "ForcedBNR_ELBM"):
        got = refbuild.run_ref(refbuild.RefConfig(collision=collision, **base), f0, 2)
        assert np.array_equal(got["f"], expected["f"]) and np.array_equal(got["alpha"], expected["alpha"]), collision
    different = refbuild.run_ref(refbuild.RefConfig(collision="ForcedNR_ELBM_Forcing", **base), f0, 2)
    assert not np.array_equal(different["alpha"], expected["alpha"])


def _reference_or_skip(ref_cfg):
    if not refbuild.reference_available() and not refbuild.binary_path(ref_cfg).is_file():
        pytest.skip("reference sources and prebuilt binary both absent")


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}-{c[2]}-{c[4]}-p{c[8]}")
def test_oracle_equals_compiled_reference(case, oracle_lib):
    lattice, shape, collision, equilibrium, scheme, force, tau, eps, ranks = case
    amplitude, wavelength = (1e-4, 2e-4, 3e-4), (8.0, 8.0, 8.0)
    ref_cfg = refbuild.RefConfig(lattice=lattice, nx=shape[0], ny=shape[1], nz=shape[2], collision=collision,
                                 equilibrium=equilibrium, forcing_scheme=scheme, force=force, tau=tau,
                                 amplitude=amplitude, wavelength=wavelength, nprocs=ranks)
    _reference_or_skip(ref_cfg)
    cfg = make_config(lattice=lattice, shape=shape, collision=collision, equilibrium=equilibrium, forcing_scheme=scheme,
                      force=force, tau=tau, amplitude=amplitude, wavelength=wavelength)
    f0 = O.synthetic_populations(cfg, eps=eps)
    reference = refbuild.run_ref(ref_cfg, f0, 3, store_every=1)
    state = O.OracleState(cfg, f0)
    for _ in range(3):
        state.step(True)
    assert np.array_equal(state.f, reference["f"])
    assert np.array_equal(state.alpha, reference["alpha"])
    energy = reference["observables"][-1][1]
    assert abs(state.observables()[0] - energy) <= 1e-13 * abs(energy)


def test_reference_native_initialisation_conserves_mass(oracle_lib):
    """Init A of SURVEY 8(d): the reference's own rho = 1, u = 0 start, Kolmogorov-forced, 20 steps."""
    ref_cfg = refbuild.RefConfig(lattice="D2Q9", nx=16, ny=16, nz=1, wavelength=(16.0, 16.0, 16.0))
    _reference_or_skip(ref_cfg)
    reference = refbuild.run_ref(ref_cfg, None, 20, store_every=0)
    assert abs(reference["density"].sum() - 256.0) <= 1e-10
    cfg = make_config(lattice="D2Q9", shape=(16, 16, 1), forcing_scheme="Guo", force="Kolmogorov", tau=0.7,
                      amplitude=(1e-5, 1e-5, 1e-5), wavelength=(16.0, 16.0, 16.0))
    f0 = O.init_equilibrium(cfg, np.ones((16, 16, 1)), np.zeros((2, 16, 16, 1)))
    state = O.OracleState(cfg, f0)
    for _ in range(20):
        state.step(True)
    assert np.array_equal(state.f, reference["f"])
