"""The CPU oracle (oracle/lbm_oracle.c) against the golden vectors the reference itself produced
(tests/golden/make_golden.py): bit-for-bit on populations, alpha, density and force."""
import numpy as np
import pytest

from golden_util import golden_names, load_golden
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
