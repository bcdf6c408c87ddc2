"""Known-answer identities the reference's (stale) unit tests pin, checked on the oracle's tables and per-node maths:
lattice isotropy (test/TestLattice.cpp:17-110), moments (test/TestMoment.cpp:27-95), forcing schemes
(test/TestForcingScheme.cpp:27-188), BGK alpha == 2 (test/TestCollision.cpp:19-27)."""
import itertools

import numpy as np
import pytest

from metalbm_b200.capi import make_config
from oracle import oracle as O

LATTICES = ["D2Q5", "D2Q9", "D3Q15", "D3Q19", "D3Q27"]
SHAPES = {2: (6, 5, 1), 3: (5, 4, 3)}


@pytest.mark.parametrize("name", LATTICES)
def test_lattice_isotropy(name, oracle_lib):
    dim, q, c, w = O.lattice(name)
    c = c[:, :dim].astype(float)
    assert abs(w.sum() - 1.0) <= 1e-15
    assert np.abs((w[:, None] * c).sum(0)).max() <= 1e-15
    second = np.einsum("q,qa,qb->ab", w, c, c)
    if name != "D2Q5":  # the reference's D2Q5 weights give cs2 = 1/6 while it declares inv_cs2 = 3 (Lattice.h:86, 135-138)
        assert np.abs(second - np.eye(dim) / 3.0).max() <= 1e-15
    assert np.abs(np.einsum("q,qa,qb,qc->abc", w, c, c, c)).max() <= 1e-15
    if name in ("D2Q9", "D3Q19", "D3Q27"):
        fourth = np.einsum("q,qa,qb,qc,qd->abcd", w, c, c, c, c)
        delta = np.eye(dim)
        expected = (np.einsum("ab,cd->abcd", delta, delta) + np.einsum("ac,bd->abcd", delta, delta)
                    + np.einsum("ad,bc->abcd", delta, delta)) / 9.0
        assert np.abs(fourth - expected).max() <= 1e-15


@pytest.mark.parametrize("name", LATTICES)
def test_halo_ordering_contract(name, oracle_lib):
    """iQ 1..faceQ have c_x < 0 and faceQ+1..2 faceQ have c_x > 0 (Communication.h:138,161)."""
    dim, q, c, w = O.lattice(name)
    face = int((c[:, 0] < 0).sum())
    assert np.all(c[1:face + 1, 0] == -1) and np.all(c[face + 1:2 * face + 1, 0] == 1) and np.all(c[2 * face + 1:, 0] == 0)


@pytest.mark.parametrize("name", ["D2Q9", "D3Q19", "D3Q27"])
def test_moments_of_equilibrium(name, oracle_lib):
    dim, q, c, w = O.lattice(name)
    shape = SHAPES[dim]
    cfg = make_config(lattice=name, shape=shape)
    rng = np.random.default_rng(3)
    rho = 1.0 + 0.1 * rng.standard_normal(shape)
    u = 0.01 * rng.standard_normal((dim,) + shape)
    f = O.init_equilibrium(cfg, rho, u)
    assert np.abs(f.sum(0) - rho).max() <= 1e-14
    momentum = np.einsum("qd,qxyz->dxyz", c[:, :dim].astype(float), f)
    assert np.abs(momentum / f.sum(0) - u).max() <= 1e-6  # the 4th-order polynomial conserves momentum to O(u^3)


def _uniform_step(name, scheme, force, amplitude=(1e-3, 2e-3, 3e-3), tau=0.8):
    dim, q, c, w = O.lattice(name)
    shape = SHAPES[dim]
    cfg = make_config(lattice=name, shape=shape, forcing_scheme=scheme, force=force, tau=tau, amplitude=amplitude)
    f0 = np.broadcast_to(w[:, None, None, None], (q,) + shape).copy()  # rho = 1, u = 0 everywhere
    state = O.OracleState(cfg, f0)
    state.step(True)
    return dim, c, w, state, f0


@pytest.mark.parametrize("name", ["D2Q9", "D3Q19"])
def test_guo_source_at_rest(name, oracle_lib):
    """Guo source = (1 - 1/(2 tau)) w inv_cs2 (c . F) at u = 0 (test/TestForcingScheme.cpp, ForcingScheme.h:99-117)."""
    tau, amplitude = 0.8, (1e-3, 2e-3, 3e-3)
    dim, c, w, state, f0 = _uniform_step(name, "Guo", "Constant", amplitude, tau)
    force = np.array(amplitude[:dim])
    expected = f0[:, 0, 0, 0] + (1 - 1 / (2 * tau)) * w * 3.0 * (c[:, :dim] @ force)
    assert np.abs(state.f[:, 0, 0, 0] - expected).max() <= 1e-15
    assert np.abs(state.velocity[:, 0, 0, 0] - 0.5 * force).max() <= 1e-15  # u_hydro = u + F / (2 rho)


@pytest.mark.parametrize("scheme", ["None", "ShanChen"])
def test_zero_source_schemes_leave_populations_unforced(scheme, oracle_lib):
    dim, c, w, state, f0 = _uniform_step("D2Q9", scheme, "Constant")
    assert np.abs(state.f - f0).max() <= 1e-15
    expected = 0.0 if scheme == "None" else 0.5e-3
    assert abs(state.velocity[0, 0, 0, 0] - expected) <= 1e-15


def test_edm_source_vanishes_without_force(oracle_lib):
    dim, c, w, state, f0 = _uniform_step("D3Q19", "ExactDifferenceMethod", "None")
    assert np.abs(state.f - f0).max() <= 1e-15


def test_bgk_alpha_is_two(oracle_lib):
    dim, c, w, state, f0 = _uniform_step("D2Q9", "Guo", "Constant")
    assert np.all(state.alpha == 2.0)


def test_kolmogorov_force_profile(oracle_lib):
    """F_x = A_x sin(y 2 pi / lambda_x), other components zero (Force.h:262-267)."""
    cfg = make_config(lattice="D3Q19", shape=(4, 8, 3), forcing_scheme="Guo", force="Kolmogorov", tau=0.8,
                      amplitude=(1e-3, 5.0, 7.0), wavelength=(8.0, 2.0, 2.0))
    dim, q, c, w = O.lattice("D3Q19")
    state = O.OracleState(cfg, np.broadcast_to(w[:, None, None, None], (q, 4, 8, 3)).copy())
    state.step(True)
    y = np.arange(8)
    assert np.array_equal(state.force[0, 0, :, 0], 1e-3 * np.sin(y * 2 * np.pi / 8.0))
    assert np.all(state.force[1:] == 0.0)
