"""Loading of the committed golden vectors (tests/golden/*.npz, produced by the reference itself)."""
import json
from pathlib import Path

import numpy as np

from metalbm_b200.capi import make_config

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"


def golden_names():
    return sorted(p.stem for p in GOLDEN_DIR.glob("*.npz"))


def load_golden(name):
    data = np.load(GOLDEN_DIR / f"{name}.npz")
    meta = json.loads(str(data["meta"]))
    cfg = make_config(lattice=meta["lattice"], shape=meta["shape"], collision=meta["collision"],
                      equilibrium=meta["equilibrium"], forcing_scheme=meta["forcing_scheme"], force=meta["force"],
                      tau=meta["tau"], amplitude=meta["amplitude"], wavelength=meta["wavelength"])
    return meta, cfg, data
