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


def run_oracle(cfg, f0, steps, alpha0=None):
    state = O.OracleState(cfg, f0, alpha0)
    for _ in range(steps):
        state.step(True)
    return state


def run_cuda(cfg, f0, steps, alpha0=None, store_last=True):
    """unpack -> iterate x steps (isStored on the last) -> pack; returns dict of global arrays (single rank)."""
    algorithm = Algorithm(cfg)
    try:
        domain = algorithm.domain
        algorithm.distribution.set_interior(f0.astype(domain.dtype))
        algorithm.unpack()
        if alpha0 is not None:
            domain.interior(algorithm.fieldList.alpha)[0] = alpha0
            algorithm.set_alpha()
        for iteration in range(1, steps + 1):
            algorithm.isStored = store_last and iteration == steps
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
