"""Loading of the committed golden vectors (tests/golden/*.npz, produced by the reference itself)."""
import json
from pathlib import Path

import numpy as np

from metalbm_b200.capi import make_config

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"


SPECTRAL_MARKERS = ("constantshell", "turbulent2d", "energyremoval")


def golden_names(spectral=None):
    """Names of the committed golden vectors; `spectral` = True / False selects / excludes the cases recorded with the
    reference's array-type spectral forces (Force.h:296-616), whose GPU tests live in tests/test_spectral_forces_gpu.py."""
    names = sorted(p.stem for p in GOLDEN_DIR.glob("*.npz"))
    if spectral is None:
        return names
    return [n for n in names if any(m in n for m in SPECTRAL_MARKERS) == spectral]


def load_golden(name):
    data = np.load(GOLDEN_DIR / f"{name}.npz")
    meta = json.loads(str(data["meta"]))
    cfg = make_config(lattice=meta["lattice"], shape=meta["shape"], collision=meta["collision"],
                      equilibrium=meta["equilibrium"], forcing_scheme=meta["forcing_scheme"], force=meta["force"],
                      tau=meta["tau"], amplitude=meta["amplitude"], wavelength=meta["wavelength"], **meta.get("shell", {}))
    return meta, cfg, data
