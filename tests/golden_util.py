"""Loading of the committed golden vectors (tests/golden/*.npz, produced by the reference itself)."""
import json
from pathlib import Path

import numpy as np

from metalbm_b200.capi import make_config

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"


def golden_names():
    """Alphabetical, the array-type / spectral force cases last (newest code paths: a failure there must not hide the rest
    of the set from a run that stops at the first failure)."""
    spectral = ("constantshell", "turbulent2d", "energyremoval")
    return sorted((p.stem for p in GOLDEN_DIR.glob("*.npz")), key=lambda name: (any(s in name for s in spectral), name))


def load_golden(name):
    data = np.load(GOLDEN_DIR / f"{name}.npz")
    meta = json.loads(str(data["meta"]))
    cfg = make_config(lattice=meta["lattice"], shape=meta["shape"], collision=meta["collision"],
                      equilibrium=meta["equilibrium"], forcing_scheme=meta["forcing_scheme"], force=meta["force"],
                      tau=meta["tau"], amplitude=meta["amplitude"], wavelength=meta["wavelength"], **meta.get("shell", {}))
    return meta, cfg, data
