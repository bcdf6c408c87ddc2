"""oracle/refbuild.py -- TEST INFRASTRUCTURE ONLY.

Compiles the reference's own CPU implementation of the fused step from the sources
where they lie under /root/reference (never copied into this repository) into
``oracle/_ref/`` and runs it.  Every physics choice of the reference is a compile-time
``constexpr`` global read from ``Input.in`` (src/Input_prod.in:10-78), so each
configuration is its own small binary; they are cached by configuration hash.

Only ``tests/``, ``__graft_entry__`` (build/smoke) and ``bench.py``'s reference / cpu_baseline
legs may use this module.  ``/root/reference`` exists only in the build container: on
the GPU box only the binaries already present in ``oracle/_ref`` can be run.
"""
from __future__ import annotations

import dataclasses
import hashlib
import json
import os
import subprocess
import tempfile
from pathlib import Path

import numpy as np

REFERENCE_ROOT = Path(os.environ.get("MLBM_REFERENCE_ROOT", "/root/reference"))
ORACLE_DIR = Path(__file__).resolve().parent
REF_DIR = ORACLE_DIR / "_ref"

LATTICES = {
    "D1Q3": (1, 3), "D2Q5": (2, 5), "D2Q9": (2, 9), "D3Q15": (3, 15), "D3Q19": (3, 19), "D3Q27": (3, 27),
    "D2Q13": (2, 13), "D2Q17": (2, 17), "D2Q21": (2, 21), "D3Q33": (3, 33),   # multi-speed (Lattice.h:213-458, 706-803)
}


@dataclasses.dataclass(frozen=True)
class RefConfig:
    """One compile-time configuration of the reference (names follow src/Input_prod.in)."""
    lattice: str = "D2Q9"
    nx: int = 16
    ny: int = 16
    nz: int = 1
    collision: str = "BGK"              # BGK | ELBM | ForcedNR_ELBM
    equilibrium: str = "TruncationMa3"  # TruncationMa3 | Exact
    forcing_scheme: str = "Guo"         # None | Guo | ShanChen | ExactDifferenceMethod
    force: str = "Kolmogorov"           # None | Kolmogorov  (Constant/Sinusoidal do not compile in the snapshot)
    tau: float = 0.7
    amplitude: tuple = (1e-5, 1e-5, 1e-5)
    wavelength: tuple = (32.0, 32.0, 32.0)
    init_density: str = "Homogeneous"   # Homogeneous | Peak
    k_min: int = 1                      # forcekMin / forcekMax: shell of the spectral forces (Force.h:296-623)
    k_max: int = 2
    removal_amplitude: tuple = (0.0, 0.0, 0.0)   # removalForce*: the EnergyRemoval half of Turbulent2D (Force.h:566-577)
    removal_k_min: int = 1
    removal_k_max: int = 2
    nprocs: int = 1
    optimize: str = "-O2"               # "-O2" = parity build (no FMA contraction); TIMING_FLAGS = timing build

    @property
    def dim(self) -> int:
        return LATTICES[self.lattice][0]

    @property
    def q(self) -> int:
        return LATTICES[self.lattice][1]

    @property
    def shape(self) -> tuple:
        return (self.nx, self.ny if self.dim > 1 else 1, self.nz if self.dim > 2 else 1)

    def key(self) -> str:
        blob = json.dumps(dataclasses.asdict(self), sort_keys=True, default=repr)
        return hashlib.sha1(blob.encode()).hexdigest()[:12]

    def name(self) -> str:
        flags = "opt" if self.optimize != "-O2" else "par"
        return (f"ref_{self.lattice}_{self.collision}_{self.equilibrium}_{self.forcing_scheme}_{self.force}"
                f"_{self.nx}x{self.ny}x{self.nz}_p{self.nprocs}_{flags}_{self.key()}")


def _vector(values) -> str:
    return "{ {" + ", ".join(repr(float(v)) for v in values) + "} }"


def input_in(cfg: RefConfig) -> str:
    """The compile-time configuration header the reference headers read as globals."""
    return f"""#pragma once
#include <string>
#include "metaLBM/Commons.h"
#include "metaLBM/Options.h"
#include "metaLBM/MathVector.h"
namespace lbm {{
  using dataT = double;
  using Vector = MathVector<dataT, 3>;
  constexpr int numProcs = NPROCS;
  constexpr int numThreads = NTHREADS;
  constexpr LatticeType latticeT = LatticeType::{cfg.lattice};
  constexpr int globalLengthX = GLOBAL_LENGTH_X;
  constexpr int globalLengthY = GLOBAL_LENGTH_Y;
  constexpr int globalLengthZ = GLOBAL_LENGTH_Z;
  constexpr unsigned int startIteration = 0;
  constexpr unsigned int endIteration = 1;
  constexpr unsigned int writeStep = 1000000000;
  constexpr unsigned int backUpStep = 1000000000;
  constexpr unsigned int scalarAnalysisStep = 1000000000;
  constexpr unsigned int spectralAnalysisStep = 1000000000;
  constexpr unsigned int performanceAnalysisStep = 1000000000;
  constexpr unsigned int successiveWriteStep = 1;
  constexpr AlgorithmType algorithmT = AlgorithmType::Pull;
  constexpr PartitionningType partitionningT = PartitionningType::OneD;
  constexpr CommunicationType communicationT = CommunicationType::MPI;
  constexpr MemoryLayout memoryL = MemoryLayout::SoA;
  constexpr Overlapping overlappingT = Overlapping::Off;
  constexpr dataT relaxationTime = {cfg.tau!r};
  constexpr CollisionType collisionT = CollisionType::{cfg.collision};
  constexpr EquilibriumType equilibriumT = EquilibriumType::{cfg.equilibrium};
  constexpr InitDensityType initDensityT = InitDensityType::{cfg.init_density};
  constexpr dataT initDensityValue = 1.0;
  constexpr InitVelocityType initVelocityT = InitVelocityType::Homogeneous;
  constexpr Vector initVelocityVector = {{ {{0.0, 0.0, 0.0}} }};
  constexpr ForcingSchemeType forcingSchemeT = ForcingSchemeType::{cfg.forcing_scheme};
  constexpr ForceType forceT = ForceType::{cfg.force};
  constexpr Vector forceAmplitude = {_vector(cfg.amplitude)};
  constexpr Vector forceWaveLength = {_vector(cfg.wavelength)};
  constexpr int forcekMin = {int(cfg.k_min)};
  constexpr int forcekMax = {int(cfg.k_max)};
  constexpr Vector removalForceAmplitude = {_vector(cfg.removal_amplitude)};
  constexpr Vector removalForceWaveLength = {{ {{32.0, 32.0, 32.0}} }};
  constexpr int removalForcekMin = {int(cfg.removal_k_min)};
  constexpr int removalForcekMax = {int(cfg.removal_k_max)};
  constexpr BoundaryType boundaryT = BoundaryType::Generic;
  constexpr InputOutputFormat inputOutputFormatT = InputOutputFormat::ascii;
  constexpr auto prefix = LBM_POSTFIX;
  constexpr bool writeFieldInit = 0;
  constexpr bool writeAnalysisInit = 0;
  constexpr bool writeForce = 1;
  constexpr bool writeEntropy = 0;
  constexpr bool writeAlpha = 1;
  constexpr bool writeT = 0;
  constexpr bool writeVorticity = 1;
  constexpr bool writeKinetics = 0;
  constexpr bool analyzeTotalEnergy = 1;
  constexpr bool analyzeTotalEnstrophy = 1;
  constexpr bool analyzeEnergySpectra = 0;
  constexpr bool analyzeEnstrophySpectra = 0;
}}
"""


def reference_available() -> bool:
    return (REFERENCE_ROOT / "include" / "metaLBM" / "Algorithm.h").is_file()


def binary_path(cfg: RefConfig) -> Path:
    return REF_DIR / cfg.name()


def build_ref(cfg: RefConfig, force: bool = False) -> Path:
    """Compile (or reuse) the reference binary for ``cfg``; returns its path."""
    out = binary_path(cfg)
    driver = ORACLE_DIR / "ref_driver.cpp"
    # a binary older than the driver is stale -- but only where it can be rebuilt (the GPU box has no reference sources)
    stale = out.is_file() and reference_available() and driver.stat().st_mtime > out.stat().st_mtime
    if out.is_file() and not force and not stale:
        return out
    if not reference_available():
        raise FileNotFoundError(f"{out} is not prebuilt and {REFERENCE_ROOT} is absent")
    REF_DIR.mkdir(parents=True, exist_ok=True)
    with tempfile.TemporaryDirectory(prefix="mlbm_ref_") as tmp:
        (Path(tmp) / "Input.in").write_text(input_in(cfg))
        nx, ny, nz = cfg.shape
        cmd = ["g++", "-std=c++14", *cfg.optimize.split(), "-ffp-contract=off" if cfg.optimize == "-O2" else "-ffp-contract=fast",
               "-w", "-DUSE_FFTW", f"-DNPROCS={cfg.nprocs}", "-DNTHREADS=1",
               f"-DGLOBAL_LENGTH_X={nx}", f"-DGLOBAL_LENGTH_Y={ny}", f"-DGLOBAL_LENGTH_Z={nz}",
               '-DLBM_POSTFIX="oracle"', f"-I{ORACLE_DIR / 'shim'}", f"-I{tmp}",
               f"-I{REFERENCE_ROOT / 'include'}", str(ORACLE_DIR / "ref_driver.cpp"), "-o", str(out) + ".tmp"]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError("reference build failed:\n" + " ".join(cmd) + "\n" + proc.stderr[-4000:])
        os.replace(str(out) + ".tmp", out)
    return out


def run_ref(cfg: RefConfig, populations: np.ndarray | None, steps: int, store_every: int = 0,
            observables: bool = True, timeout: float = 3600.0, dump: bool = True, warmup: int = 0) -> dict:
    """Run the reference binary.  ``populations`` is float64 [Q, nx, ny, nz] (global interior) or None
    for the reference's own equilibrium initialisation.  Returns the final populations / fields and
    the (iteration, energy, enstrophy) rows of every stored step."""
    exe = build_ref(cfg)
    nx, ny, nz = cfg.shape
    q, dim = cfg.q, cfg.dim
    with tempfile.TemporaryDirectory(prefix="mlbm_run_") as tmp:
        tmp = Path(tmp)
        if populations is None:
            inp = "-"
        else:
            arr = np.ascontiguousarray(populations, dtype=np.float64).reshape(q, nx, ny, nz)
            inp = str(tmp / "f0.bin")
            arr.tofile(inp)
        prefix = str(tmp / "out")
        proc = subprocess.run([str(exe), inp, prefix, str(steps), str(store_every), "1" if observables else "0",
                               "1" if dump else "0", str(warmup)],
                              capture_output=True, text=True, timeout=timeout, cwd=tmp)
        if proc.returncode != 0:
            raise RuntimeError(f"reference run failed ({proc.returncode}):\n{proc.stdout[-2000:]}\n{proc.stderr[-2000:]}")
        result = {"observables": []}
        for line in Path(prefix + ".txt").read_text().splitlines():
            parts = line.split()
            if parts[0] == "obs":
                result["observables"].append((int(parts[1]), float(parts[2]), float(parts[3])))
            elif parts[0] == "spec":   # wave number, energy spectrum, forcing spectrum of the last step's fields
                result.setdefault("spectra", []).append((float(parts[2]), float(parts[3])))
            elif parts[0].startswith("time_"):
                result[parts[0]] = float(parts[1])
        if not dump:
            return result
        lx = nx // cfg.nprocs
        nvort = 2 * dim - 3
        nfields = q + 1 + dim + 1 + dim + nvort
        slabs = [np.fromfile(f"{prefix}.r{r}.bin", dtype=np.float64).reshape(nfields, lx, ny, nz)
                 for r in range(cfg.nprocs)]
        full = np.concatenate(slabs, axis=1)
        result.update({
            "f": full[:q].copy(),
            "density": full[q].copy(),
            "velocity": full[q + 1:q + 1 + dim].copy(),
            "alpha": full[q + 1 + dim].copy(),
            "force": full[q + 2 + dim:q + 2 + 2 * dim].copy(),
            "vorticity": full[q + 2 + 2 * dim:].copy(),
        })
        return result


# The reference's performance build.  x86-64-v3 (AVX2 + FMA) rather than -march=native because the binary is
# compiled in the build container and timed on the GPU box, whose host CPU may differ.
TIMING_FLAGS = "-O3 -march=x86-64-v3 -funroll-loops"
TIMING_RANKS = (1, 2, 4, 8, 16, 32, 64, 128)
TIMING_PLANES_PER_RANK = 2   # the thin-slab sample of round 1 (kept as a second figure)
TIMING_EDGE = 256


LARGE_EDGE = 1024            # cross-section of BASELINE configs[4] (D3Q19 1024^3, the multi-GPU headline)
LARGE_MAX_RANKS = 32         # 2 planes of 1024^2 per rank: 3 arrays x 19 x 8 B x 2.1e6 nodes = 0.96 GB per rank


def timing_config(nprocs: int, planes: int | None = None, edge: int = TIMING_EDGE) -> RefConfig:
    """The headline workload (D3Q19 SRT-BGK, edge x edge cross-section, FP64).  planes=None: the WHOLE cube of
    BASELINE configs[1] (edge 256) split into nprocs x-slabs, one rank per host core (256 / nprocs planes per rank) -- the
    same configuration the GPU arm times.  planes=k: a bounded x-slab sample of k planes per rank (k = 2 makes the
    reference's halo exchange and y/z boundary copies as expensive as its node update); edge=1024 samples configs[4]."""
    nx = edge if planes is None else planes * nprocs
    return RefConfig(lattice="D3Q19", nx=nx, ny=edge, nz=edge, collision="BGK",
                     forcing_scheme="None", force="None", tau=0.55, nprocs=nprocs, optimize=TIMING_FLAGS)


def timing_configs() -> list:
    """Prebuilt by __graft_entry__.build() so that they travel to the GPU box (one binary per rank count,
    numProcs is a compile-time constant of the reference): the full 256^3 cube, its thin-slab sample, and the thin-slab
    sample of the 1024^3 cube."""
    configs = []
    for n in TIMING_RANKS:
        candidates = [timing_config(n, None), timing_config(n, TIMING_PLANES_PER_RANK)]
        if n <= LARGE_MAX_RANKS:
            candidates.append(timing_config(n, TIMING_PLANES_PER_RANK, LARGE_EDGE))
        for cfg in candidates:
            if all(cfg.name() != other.name() for other in configs):
                configs.append(cfg)
    return configs


def best_timing_config(cores: int, planes: int | None = None, edge: int = TIMING_EDGE) -> RefConfig | None:
    """Largest prebuilt rank count that fits the host cores."""
    for n in sorted(TIMING_RANKS, reverse=True):
        if n <= cores and binary_path(timing_config(n, planes, edge)).is_file():
            return timing_config(n, planes, edge)
    return None


def time_reference(cfg: RefConfig, steps: int, warmup: int = 0, timeout: float = 3600.0) -> dict:
    """MLUPS of the reference CPU build on cfg (its own initialisation, no observables) over `steps` timed
    steps after `warmup` untimed ones."""
    result = run_ref(cfg, None, steps + warmup, 0, observables=False, timeout=timeout, dump=False, warmup=warmup)
    nodes = cfg.nx * cfg.ny * cfg.nz
    # per-rank average of the reference's own timers (Algorithm.h:340-357) = time of the parallel run
    seconds = result["time_computation"] + result["time_communication"]
    return {"mlups": nodes * steps / seconds / 1e6, "seconds": seconds, "wall": result["time_wall"],
            "nodes": nodes, "steps": steps, "ranks": cfg.nprocs,
            "communication_share": result["time_communication"] / seconds if seconds > 0 else None}


if __name__ == "__main__":
    import sys
    cfg = RefConfig()
    print(build_ref(cfg, force="--force" in sys.argv))
