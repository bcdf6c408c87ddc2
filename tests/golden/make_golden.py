"""Generates the golden vectors in this directory by RUNNING THE REFERENCE ITSELF (oracle/_ref binaries compiled
from /root/reference by oracle/refbuild.py, parity flags: -O2 -ffp-contract=off).  Only runs in the build
container; the .npz files it writes are committed so that the GPU box (which has no /root/reference) can check
the CUDA path and the C oracle against real reference outputs.

    python tests/golden/make_golden.py
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from metalbm_b200.capi import make_config  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle.refbuild import RefConfig, run_ref  # noqa: E402

HERE = Path(__file__).resolve().parent

# name, lattice, shape, collision, equilibrium, scheme, force, tau, eps, flow amplitude, density ripple, steps, ranks
CASES = [
    ("d2q9_bgk_guo_kolmogorov", "D2Q9", (24, 20, 1), "BGK", "TruncationMa3", "Guo", "Kolmogorov", 0.7, 1e-2, 0.05, 0.05, 3, 1),
    ("d2q9_bgk_guo_kolmogorov_100", "D2Q9", (24, 20, 1), "BGK", "TruncationMa3", "Guo", "Kolmogorov", 0.7, 1e-2, 0.05, 0.05, 100, 1),
    ("d2q9_bgk_none", "D2Q9", (10, 14, 1), "BGK", "TruncationMa3", "None", "None", 0.6, 1e-2, 0.05, 0.05, 2, 1),
    ("d2q9_bgk_exact_edm", "D2Q9", (12, 10, 1), "BGK", "Exact", "ExactDifferenceMethod", "Kolmogorov", 0.7, 1e-2, 0.05, 0.05, 2, 1),
    ("d2q9_elbm_shanchen", "D2Q9", (16, 12, 1), "ELBM", "TruncationMa3", "ShanChen", "Kolmogorov", 0.51, 2e-2, 0.05, 0.05, 2, 1),
    ("d2q9_elbm_edm", "D2Q9", (16, 12, 1), "ELBM", "TruncationMa3", "ExactDifferenceMethod", "Kolmogorov", 0.51, 2e-2, 0.05, 0.05, 2, 1),
    ("d2q9_elbm_small_deviation", "D2Q9", (16, 12, 1), "ELBM", "TruncationMa3", "Guo", "Kolmogorov", 0.55, 4e-4, 0.0, 0.0, 2, 1),
    ("d2q5_bgk_guo", "D2Q5", (8, 10, 1), "BGK", "TruncationMa3", "Guo", "Kolmogorov", 0.8, 1e-2, 0.05, 0.05, 2, 1),
    ("d3q15_bgk_edm", "D3Q15", (6, 8, 4), "BGK", "TruncationMa3", "ExactDifferenceMethod", "Kolmogorov", 0.6, 1e-2, 0.05, 0.05, 2, 1),
    ("d3q19_bgk_none", "D3Q19", (8, 6, 10), "BGK", "TruncationMa3", "None", "None", 0.55, 1e-3, 0.05, 0.05, 3, 1),
    ("d3q19_bgk_guo_kolmogorov", "D3Q19", (8, 6, 4), "BGK", "TruncationMa3", "Guo", "Kolmogorov", 0.55, 1e-3, 0.05, 0.05, 3, 1),
    ("d3q19_bgk_guo_2ranks", "D3Q19", (8, 6, 4), "BGK", "TruncationMa3", "Guo", "Kolmogorov", 0.55, 1e-3, 0.05, 0.05, 3, 2),
    ("d3q19_elbm_all_branches", "D3Q19", (8, 6, 4), "ELBM", "TruncationMa3", "Guo", "Kolmogorov", 0.55, 3e-1, 0.05, 0.05, 2, 1),
    ("d3q27_elbm_guo", "D3Q27", (8, 6, 4), "ELBM", "TruncationMa3", "Guo", "Kolmogorov", 0.55, 2e-2, 0.05, 0.05, 2, 1),
    ("d3q27_elbm_exact", "D3Q27", (8, 6, 4), "ELBM", "Exact", "Guo", "Kolmogorov", 0.50000032, 2e-2, 0.05, 0.05, 2, 1),
    ("d3q27_forcednr_elbm", "D3Q27", (6, 6, 4), "ForcedNR_ELBM", "TruncationMa3", "Guo", "Kolmogorov", 0.55, 2e-2, 0.05, 0.05, 2, 1),
    # the remaining alpha models of Collision.h: dead overrides, the reference runs them exactly like ELBM
    ("d2q9_approached_elbm", "D2Q9", (12, 10, 1), "Approached_ELBM", "TruncationMa3", "Guo", "Kolmogorov", 0.51, 2e-2, 0.05, 0.05, 2, 1),
    ("d2q9_malaspinas_elbm", "D2Q9", (12, 10, 1), "Malaspinas_ELBM", "TruncationMa3", "ExactDifferenceMethod", "Kolmogorov", 0.51, 2e-2, 0.05, 0.05, 2, 1),
    ("d3q19_essentially1_elbm", "D3Q19", (6, 4, 4), "Essentially1_ELBM", "TruncationMa3", "Guo", "Kolmogorov", 0.55, 3e-1, 0.05, 0.05, 2, 1),
    ("d3q19_essentially2_elbm", "D3Q19", (6, 4, 4), "Essentially2_ELBM", "TruncationMa3", "ShanChen", "Kolmogorov", 0.55, 2e-2, 0.05, 0.05, 2, 1),
    ("d3q27_forcedbnr_elbm", "D3Q27", (6, 6, 4), "ForcedBNR_ELBM", "TruncationMa3", "Guo", "Kolmogorov", 0.55, 2e-2, 0.05, 0.05, 2, 1),
    # ForcedNR_ELBM_Forcing (Collision.h:727-857): alpha solved on the forced populations (oracle only, no device kernel yet)
    ("d2q9_forcednr_elbm_forcing", "D2Q9", (12, 10, 1), "ForcedNR_ELBM_Forcing", "TruncationMa3", "Guo", "Kolmogorov", 0.51, 2e-2, 0.05, 0.05, 3, 1),
    ("d3q19_forcednr_elbm_forcing", "D3Q19", (6, 4, 4), "ForcedNR_ELBM_Forcing", "TruncationMa3", "Guo", "Kolmogorov", 0.55, 3e-1, 0.05, 0.05, 2, 1),
    ("d3q27_forcednr_elbm_forcing_edm", "D3Q27", (6, 6, 4), "ForcedNR_ELBM_Forcing", "TruncationMa3", "ExactDifferenceMethod", "Kolmogorov", 0.55, 2e-2, 0.05, 0.05, 2, 1),
    # multi-speed lattices (Lattice.h:213-458, 706-803): jumps of up to 3 nodes, their own sound speeds
    ("d2q13_bgk_guo", "D2Q13", (14, 10, 1), "BGK", "TruncationMa3", "Guo", "Kolmogorov", 0.7, 1e-2, 0.05, 0.05, 3, 1),
    ("d2q17_bgk_edm", "D2Q17", (12, 10, 1), "BGK", "TruncationMa3", "ExactDifferenceMethod", "Kolmogorov", 0.7, 1e-2, 0.05, 0.05, 3, 1),
    ("d2q21_elbm_guo", "D2Q21", (12, 10, 1), "ELBM", "TruncationMa3", "Guo", "Kolmogorov", 0.55, 2e-2, 0.05, 0.05, 2, 1),
    ("d3q33_bgk_guo", "D3Q33", (6, 5, 4), "BGK", "TruncationMa3", "Guo", "Kolmogorov", 0.6, 1e-2, 0.05, 0.05, 3, 1),
    ("d3q33_elbm_shanchen", "D3Q33", (6, 5, 4), "ELBM", "TruncationMa3", "ShanChen", "Kolmogorov", 0.55, 2e-2, 0.05, 0.05, 2, 1),
    # array-type forces (Force.h:296-623): the reference fills fieldList.force spectrally (its FFTs run on the oracle's DFT
    # stub of FFTW) and the step reads it through Force<Generic>::setForce (Force.h:39-48).  The golden file carries that
    # array; this repository's configuration is force "Field" fed with it
    ("d2q9_bgk_guo_constantshell", "D2Q9", (16, 12, 1), "BGK", "TruncationMa3", "Guo", "ConstantShell", 0.7, 1e-2, 0.05, 0.05, 3, 1),
    ("d2q9_elbm_edm_constantshell", "D2Q9", (16, 12, 1), "ELBM", "TruncationMa3", "ExactDifferenceMethod", "ConstantShell", 0.51, 2e-2, 0.05, 0.05, 2, 1),
    ("d2q9_bgk_shanchen_turbulent2d", "D2Q9", (10, 14, 1), "BGK", "TruncationMa3", "ShanChen", "Turbulent2D", 0.6, 1e-2, 0.05, 0.05, 3, 1),
    # (in 3-D the reference's spectral forces corrupt the heap under the single-rank shim -- Force.h:341-355 writes the
    # mirrored index of a padded local array -- so the 3-D array read is pinned by the oracle and by identities only)
]
ARRAY_FORCES = {"ConstantShell", "EnergyRemoval", "Turbulent2D"}
# time-dependent spectral forces (Force.h:423-616): the force array changes with the stored fields, so these are replayed
# natively (force "EnergyRemoval" / "Turbulent2D" of this repository), not through a recorded array
SPECTRAL_CASES = [
    dict(name="d2q9_bgk_guo_energyremoval", lattice="D2Q9", shape=(16, 12, 1), collision="BGK", scheme="Guo", force="EnergyRemoval",
         tau=0.7, eps=1e-2, steps=4, amplitude=(2e-3, 3e-3, 0.0), k_min=1, k_max=3),
    dict(name="d2q9_elbm_guo_turbulent2d_removal", lattice="D2Q9", shape=(16, 12, 1), collision="ELBM", scheme="Guo", force="Turbulent2D",
         tau=0.51, eps=2e-2, steps=3, amplitude=(1e-4, 0.0, 0.0), k_min=1, k_max=2, removal_amplitude=(2e-3, 3e-3, 0.0),
         removal_k_min=2, removal_k_max=4),
    dict(name="d2q9_bgk_edm_turbulent2d_removal_odd", lattice="D2Q9", shape=(15, 10, 1), collision="BGK", scheme="ExactDifferenceMethod",
         force="Turbulent2D", tau=0.6, eps=1e-2, steps=4, amplitude=(2e-4, 0.0, 0.0), k_min=2, k_max=3,
         removal_amplitude=(5e-3, 1e-3, 0.0), removal_k_min=0, removal_k_max=5),
]
ONLY = set(sys.argv[1:])   # optional: names of the cases to (re)generate; default all
AMPLITUDE = (1e-4, 2e-4, 3e-4)
WAVELENGTH = (8.0, 4.0, 16.0)


def _spectra(out):
    """[K, 2] energy / forcing spectra of the last step's stored fields (SpectralAnalysisList, AnalysisList.h:99-202);
    single-rank runs only (the FFT stub is single-rank)."""
    return {"spectra": np.array(out["spectra"], dtype=np.float64)} if "spectra" in out else {}


def main():
    for name, lattice, shape, collision, equilibrium, scheme, force, tau, eps, flow, ripple, steps, ranks in CASES:
        if ONLY and name not in ONLY:
            continue
        ref_cfg = RefConfig(lattice=lattice, nx=shape[0], ny=shape[1], nz=shape[2], collision=collision,
                            equilibrium=equilibrium, forcing_scheme=scheme, force=force, tau=tau,
                            amplitude=AMPLITUDE, wavelength=WAVELENGTH, nprocs=ranks)
        reference_force, force = force, ("Field" if force in ARRAY_FORCES else force)
        cfg = make_config(lattice=lattice, shape=shape, collision=collision, equilibrium=equilibrium,
                          forcing_scheme=scheme, force=force, tau=tau, amplitude=AMPLITUDE, wavelength=WAVELENGTH)
        f0 = O.synthetic_populations(cfg, eps=eps, amplitude=flow, ripple=ripple)
        out = run_ref(ref_cfg, f0, steps, store_every=1)
        meta = dict(name=name, lattice=lattice, shape=list(shape), collision=collision, equilibrium=equilibrium,
                    forcing_scheme=scheme, force=force, tau=tau, amplitude=list(AMPLITUDE), wavelength=list(WAVELENGTH),
                    steps=steps, ranks=ranks, eps=eps, reference_force=reference_force,
                    source="oracle/_ref (unmodified reference, g++ -O2 -ffp-contract=off), oracle/ref_driver.cpp")
        np.savez_compressed(HERE / f"{name}.npz", meta=json.dumps(meta), f0=f0, f=out["f"], alpha=out["alpha"],
                            density=out["density"], velocity=out["velocity"], force=out["force"],
                            observables=np.array(out["observables"], dtype=np.float64), **_spectra(out))
        print(name, out["f"].shape, out["observables"][-1])


def spectral():
    for case in SPECTRAL_CASES:
        if ONLY and case["name"] not in ONLY:
            continue
        shape = case["shape"]
        shell = {k: case[k] for k in ("k_min", "k_max", "removal_amplitude", "removal_k_min", "removal_k_max") if k in case}
        ref_cfg = RefConfig(lattice=case["lattice"], nx=shape[0], ny=shape[1], nz=shape[2], collision=case["collision"],
                            forcing_scheme=case["scheme"], force=case["force"], tau=case["tau"], amplitude=case["amplitude"],
                            wavelength=WAVELENGTH, **shell)
        cfg = make_config(lattice=case["lattice"], shape=shape, collision=case["collision"], forcing_scheme=case["scheme"],
                          tau=case["tau"])
        f0 = O.synthetic_populations(cfg, eps=case["eps"], amplitude=0.05, ripple=0.05)
        out = run_ref(ref_cfg, f0, case["steps"], store_every=1)
        meta = dict(name=case["name"], lattice=case["lattice"], shape=list(shape), collision=case["collision"],
                    equilibrium="TruncationMa3", forcing_scheme=case["scheme"], force=case["force"], tau=case["tau"],
                    amplitude=list(case["amplitude"]), wavelength=list(WAVELENGTH), steps=case["steps"], ranks=1, eps=case["eps"],
                    reference_force=case["force"], native_spectral=True, shell={k: (list(v) if isinstance(v, tuple) else v) for k, v in shell.items()},
                    source="oracle/_ref (unmodified reference, g++ -O2 -ffp-contract=off), oracle/ref_driver.cpp")
        np.savez_compressed(HERE / f"{case['name']}.npz", meta=json.dumps(meta), f0=f0, f=out["f"], alpha=out["alpha"],
                            density=out["density"], velocity=out["velocity"], force=out["force"],
                            observables=np.array(out["observables"], dtype=np.float64), **_spectra(out))
        print(case["name"], out["f"].shape, out["observables"][-1], "max |F|", np.abs(out["force"]).max())


if __name__ == "__main__":
    main()
    spectral()
