"""GPU parity of the multi-speed lattices D2Q13, D2Q17, D2Q21 and D3Q33 (Lattice.h:213-458, 706-803): jumps of
up to three nodes that wrap by index arithmetic in all three axes, sound speeds that differ from 1/3, entropic speed classes up
to |c|^2 = 18.  Against golden vectors of the reference and against the oracle (bit-identical to the compiled reference on
these lattices, tests/test_oracle_vs_reference.py).  In a file of its own that sorts after the established parity suites."""
import numpy as np
import pytest

from golden_util import golden_names
from helpers import check_entropic, relative_error, run_cuda, run_oracle
from metalbm_b200.capi import make_config
from oracle import oracle as O
from test_cpp_shim import check_template_api
from test_golden_gpu import check_cuda_against_golden
from test_multi_gpu import _device_count, _run_ranks

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_names(wide=True))
def test_cuda_reproduces_reference_outputs_on_multi_speed_lattices(name):
    check_cuda_against_golden(name)


CASES = [
    # lattice, shape, collision, scheme, force, tau, eps, dtype
    ("D2Q13", (33, 130, 1), "BGK", "Guo", "Kolmogorov", 0.7, 1e-2, "F64"),
    ("D2Q17", (16, 12, 1), "BGK", "ExactDifferenceMethod", "Sinusoidal", 0.7, 1e-2, "F64"),
    ("D2Q21", (2, 4, 1), "BGK", "ShanChen", "Kolmogorov", 0.7, 1e-2, "F64"),          # extents below the longest jump
    ("D3Q33", (10, 6, 9), "BGK", "Guo", "Constant", 0.6, 1e-2, "F64"),
    ("D2Q21", (24, 20, 1), "BGK", "Guo", "Kolmogorov", 0.7, 1e-2, "F32"),
    ("D2Q13", (16, 140, 1), "ELBM", "Guo", "Kolmogorov", 0.55, 2e-2, "F64"),
    ("D2Q17", (16, 12, 1), "ELBM", "ExactDifferenceMethod", "Kolmogorov", 0.55, 2e-2, "F64"),
    ("D2Q21", (16, 12, 1), "ForcedNR_ELBM_Forcing", "Guo", "Kolmogorov", 0.55, 2e-2, "F64"),
    ("D3Q33", (8, 6, 4), "ELBM", "Guo", "Kolmogorov", 0.55, 2e-2, "F64"),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(map(str, (c[0], "x".join(map(str, c[1])), c[2], c[3], c[7]))))
def test_multi_speed_lattices_against_the_oracle(case):
    lattice, shape, collision, scheme, force, tau, eps, dtype = case
    cfg = make_config(lattice=lattice, shape=shape, collision=collision, forcing_scheme=scheme, force=force, tau=tau,
                      amplitude=(1e-4, 2e-4, 3e-4), wavelength=(8.0, 4.0, 16.0), dtype=dtype)
    f0 = O.synthetic_populations(cfg, eps=eps)
    for steps in (1, 3):
        got = run_cuda(cfg, f0, steps)
        ref = run_oracle(cfg, f0, steps)
        if dtype == "F32":
            assert relative_error(got["f"], ref.f) <= 1e-5
            continue
        if collision == "BGK":
            assert relative_error(got["f"], ref.f) <= 1e-12
            assert relative_error(got["density"], ref.density) <= 1e-12
            assert np.abs(got["velocity"] - ref.velocity).max() <= 1e-13
        else:
            check_entropic(got, ref, cfg, steps, mismatch_budget=1e-3)
        obs = ref.observables()
        assert abs(got["observables"][0] - obs[0]) <= 1e-9 * abs(obs[0])
        assert abs(got["observables"][2] - obs[2]) <= 1e-12 * abs(obs[2]) + (0 if collision == "BGK" else 1e-9)   # Mach with the lattice's c_s
        assert abs(got["observables"][3] - obs[3]) <= 1e-12 * abs(obs[3]) + (0 if collision == "BGK" else 1e-9)


@pytest.mark.parametrize("case", [("D2Q13", (24, 20, 1), "BGK", "TruncationMa3", "Guo", "Kolmogorov", 3),
                                  ("D3Q33", (8, 6, 4), "ELBM", "TruncationMa3", "Guo", "Kolmogorov", 1)],
                         ids=lambda c: "-".join(map(str, c[:1] + c[2:6])))
def test_template_api_on_multi_speed_lattices(tmp_path, cuda_lib, case):
    """latticeT = D2Q13 / D3Q33 through the reference's template spellings (one rank)."""
    check_template_api(tmp_path, 1, case)


@pytest.mark.parametrize("overlap", ["On", "Off"])
@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("lattice,shape,collision", [("D2Q13", (16, 12, 1), "BGK"), ("D2Q17", (24, 12, 1), "ELBM"), ("D3Q33", (16, 5, 4), "BGK")])
def test_multi_speed_lattices_on_slabs(tmp_path, world, lattice, shape, collision, overlap):
    """x-slabs with dimH halo planes per side, exchanged over NCCL before the step (Off) or overlapped with the bulk planes
    (On: the first and last dimH planes first).  Bit-identical to the single-GPU run."""
    if _device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    steps = 3
    config = dict(lattice=lattice, shape=list(shape), collision=collision, forcing_scheme="Guo", force="Kolmogorov", tau=0.6,
                  amplitude=[1e-4, 2e-4, 3e-4], wavelength=[8.0, 4.0, 16.0], overlap=overlap)
    single = make_config(**config)
    f0 = O.synthetic_populations(single, eps=1e-2)
    got = _run_ranks(tmp_path, world, config, f0, steps, "sync", False)
    one = run_cuda(single, f0, steps)
    assert np.array_equal(got["f"], one["f"]) and np.array_equal(got["alpha"], one["alpha"])
    ref = run_oracle(single, f0, steps)
    if collision == "BGK":
        assert relative_error(got["f"], ref.f) <= 1e-12 * steps
    assert abs(got["observables"][0][0] - ref.observables()[0]) <= 1e-9 * abs(ref.observables()[0])
