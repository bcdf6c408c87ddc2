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
