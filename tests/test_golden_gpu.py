"""The CUDA path against the golden vectors produced by the reference itself (tests/golden/*.npz)."""
import numpy as np
import pytest

from golden_util import golden_names, load_golden
from helpers import check_entropic, relative_error, run_cuda, run_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_names(spectral=False, wide=False))
def test_cuda_reproduces_reference_outputs(name):
    check_cuda_against_golden(name)


def check_cuda_against_golden(name):
    meta, cfg, data = load_golden(name)
    force = data["force"] if meta["force"] == "Field" else None   # the array the reference's spectral force filled
    native_spectral = bool(meta.get("native_spectral"))   # EnergyRemoval / Turbulent2D follow the stored fields: the
    got = run_cuda(cfg, data["f0"], meta["steps"], force=force, store_every_step=native_spectral)   # reference stored every step
    entropic = meta["collision"] != "BGK"
    if entropic:
        # the oracle (bit-identical to the reference: tests/test_oracle_golden.py) supplies the conditioning of the
        # Newton solve at every node; the values compared are the reference's own (the golden file)
        ref = run_oracle(cfg, data["f0"], meta["steps"], force=force)
        if native_spectral:   # numpy's transforms and the reference's round differently (tests/test_oracle_golden.py)
            assert np.abs(ref.alpha - data["alpha"]).max() <= 1e-10 and relative_error(ref.f, data["f"]) <= 1e-14
        else:
            assert np.array_equal(ref.alpha, data["alpha"]) and np.array_equal(ref.f, data["f"])
        _, population_tolerance = check_entropic(got, ref, cfg, meta["steps"], mismatch_budget=5e-3)
        # density = sum of Q populations that each carry the alpha-inherited uncertainty of the previous step
        density_tolerance = max(1e-12, ref.q * population_tolerance / np.abs(data["density"]).max())
    else:
        tolerance = 1e-12 if meta["steps"] <= 3 else 1e-11
        assert relative_error(got["f"], data["f"]) <= tolerance
        assert np.all(got["alpha"] == 2.0)
        density_tolerance = 1e-12
    assert relative_error(got["density"], data["density"]) <= density_tolerance
    if native_spectral:   # EnergyRemoval / Turbulent2D: mode sums on the device, FFTs in the reference
        assert np.abs(got["force"] - data["force"]).max() <= 1e-12 * np.abs(data["force"]).max()
    else:
        assert np.array_equal(got["force"], data["force"])
    energy = data["observables"][-1][1]
    assert abs(got["observables"][0] - energy) <= 1e-9 * abs(energy)
    if meta["ranks"] == 1 and not entropic:   # the reference's own TotalEnstrophy (its Curl ran on the oracle's DFT stub)
        enstrophy = data["observables"][-1][2]
        assert abs(got["observables"][1] - enstrophy) <= 1e-9 * abs(enstrophy)
