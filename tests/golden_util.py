"""Loading of the committed golden vectors (tests/golden/*.npz, produced by the reference itself)."""
import json
from pathlib import Path

import numpy as np

from metalbm_b200.capi import make_config

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"


SPECTRAL_MARKERS = ("constantshell", "turbulent2d", "energyremoval")
WIDE_MARKERS = ("d2q13", "d2q17", "d2q21", "d3q33")   # multi-speed lattices: GPU tests in tests/test_wide_lattices_gpu.py


def golden_names(spectral=None, wide=None):
    """Names of the committed golden vectors; `spectral` / `wide` = True / False select / exclude the cases recorded with the
    reference's array-type spectral forces (Force.h:296-616) and with its multi-speed lattices, whose GPU tests live in files
    of their own (tests/test_spectral_forces_gpu.py, tests/test_wide_lattices_gpu.py)."""
    names = sorted(p.stem for p in GOLDEN_DIR.glob("*.npz"))
    if spectral is not None:
        names = [n for n in names if any(m in n for m in SPECTRAL_MARKERS) == spectral]
    if wide is not None:
        names = [n for n in names if any(m in n for m in WIDE_MARKERS) == wide]
    return names


def load_golden(name):
    data = np.load(GOLDEN_DIR / f"{name}.npz")
    meta = json.loads(str(data["meta"]))
    cfg = make_config(lattice=meta["lattice"], shape=meta["shape"], collision=meta["collision"],
                      equilibrium=meta["equilibrium"], forcing_scheme=meta["forcing_scheme"], force=meta["force"],
                      tau=meta["tau"], amplitude=meta["amplitude"], wavelength=meta["wavelength"], **meta.get("shell", {}))
    return meta, cfg, data
