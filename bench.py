#!/usr/bin/env python
"""bench.py -- MLUPS of the fused collide-and-stream pull step on B200 (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU implementation

N = 1 workload: BASELINE.json configs[1], D3Q19 SRT-BGK periodic 256^3 FP64 (one step = one pass of the
fused kernel over all 256^3 nodes; the two 2.6 GB population buffers are far larger than the 126 MB L2,
so every timed step streams from HBM -- no L2 flush needed).  N > 1 (launched with torchrun, one rank per
GPU): BASELINE.json configs[4], D3Q19 SRT-BGK 1024^3 STRONG-scaled over the N GPUs (x-slabs of 1024 / N planes),
halo exchange by direct peer stores overlapped with the bulk kernel, energy / enstrophy / Mach reductions every 50 steps
(the reductions alone where the field arrays do not fit next to the slab: 2 GPUs).  The driver's efficiency
v_N / (N v_1) is then the parallel efficiency the north star quotes its 85 % target on.

With the default headline workload the same line also carries the other BASELINE configs at this GPU count under "also"
(device-resident throughput, kernel time, both roofline fractions; N > 1: the weak-scaled 256^3-per-GPU cube, 512^3 and the
entropic configs strong-scaled, 1024^3 with whole fields), measured after the headline is final and under a watchdog
(`run_secondary`), and a one-string-per-row "also_summary" as the line's LAST key; `--also off` skips them.

One JSON line on stdout from rank 0 (see the keys at the bottom).  `value` is device-timed with inputs
resident in HBM; `e2e` is the same metric through the public C-ABI with HOST buffers: the timed region
uploads the populations from pinned host memory (Algorithm::unpack), runs the K steps reading the all-reduced scalar
observables back to the host after every step, and downloads the populations (Algorithm::pack).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "MLUPS (D3Q19, FP64)"
UNIT = "MLUPS"
LATTICE, Q, EDGE = "D3Q19", 19, 256
BYTES_PER_NODE = 2 * Q * 8           # algorithmic HBM traffic per node-step: read + write every population once
FALLBACK_PEAK_GBS = 6650.0           # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent

# BASELINE.json configs as named workloads.  `shape` is per GPU for weak scaling and global for strong scaling.
# The default (and the only one the reference arm / cpu_baseline leg time) is configs[1].
WORKLOADS = {
    "d3q19_bgk_256": dict(config=1, lattice="D3Q19", q=19, shape=(256, 256, 256), scaling="weak", collision="BGK",
                          equilibrium="TruncationMa3", scheme="None", force="None", tau=0.55, eps=0.0, store_every=0,
                          text="D3Q19 SRT-BGK periodic 256^3 FP64 per GPU (BASELINE configs[1])"),
    "d3q19_bgk_guo_256": dict(config=1, lattice="D3Q19", q=19, shape=(256, 256, 256), scaling="weak", collision="BGK",
                              equilibrium="TruncationMa3", scheme="Guo", force="Kolmogorov", tau=0.55, eps=0.0, store_every=0,
                              text="D3Q19 SRT-BGK Guo-forced Kolmogorov periodic 256^3 per GPU (second row of BASELINE configs[1])"),
    "d3q27_bgk_512": dict(config=2, lattice="D3Q27", q=27, shape=(512, 512, 512), scaling="strong", collision="BGK",
                          equilibrium="TruncationMa3", scheme="Guo", force="Kolmogorov", tau=0.55, eps=0.0, store_every=0,
                          text="D3Q27 SRT-BGK Guo Kolmogorov 512^3 (the HBM-bound counterpart of BASELINE configs[2])"),
    "d3q27_elbm_512": dict(config=2, lattice="D3Q27", q=27, shape=(512, 512, 512), scaling="strong", collision="ELBM",
                           equilibrium="TruncationMa3", scheme="Guo", force="Kolmogorov", tau=0.55, eps=2e-2, store_every=0,
                           text="D3Q27 SRT-Entropic (alpha Newton solve) Guo Kolmogorov 512^3, x-slab (BASELINE configs[2])"),
    "d3q27_elbm_512_resolved": dict(config=2, lattice="D3Q27", q=27, shape=(512, 512, 512), scaling="strong", collision="ELBM",
                                    equilibrium="TruncationMa3", scheme="Guo", force="Kolmogorov", tau=0.55, eps=1e-6, store_every=0,
                                    flow=0.005,
                                    text="D3Q27 SRT-Entropic Guo Kolmogorov 512^3 on a RESOLVED flow (velocity and density amplitude "
                                         "0.005: every node stays on the alpha = 2 shortcut, |fNeq| / f < 1e-3) -- the shortcut regime "
                                         "of BASELINE configs[2]"),
    "d2q9_elbm_shanchen_8192": dict(config=3, lattice="D2Q9", q=9, shape=(8192, 8192, 1), scaling="strong", collision="ELBM",
                                    equilibrium="TruncationMa3", scheme="ShanChen", force="Kolmogorov", tau=0.55, eps=2e-2,
                                    store_every=0, text="D2Q9 SRT-Entropic Shan-Chen Kolmogorov 8192^2 (BASELINE configs[3])"),
    "d2q9_elbm_edm_8192": dict(config=3, lattice="D2Q9", q=9, shape=(8192, 8192, 1), scaling="strong", collision="ELBM",
                               equilibrium="TruncationMa3", scheme="ExactDifferenceMethod", force="Kolmogorov", tau=0.55,
                               eps=2e-2, store_every=0, text="D2Q9 SRT-Entropic EDM Kolmogorov 8192^2 (BASELINE configs[3])"),
    "d3q19_bgk_1024": dict(config=4, lattice="D3Q19", q=19, shape=(1024, 1024, 1024), scaling="strong", collision="BGK",
                           equilibrium="TruncationMa3", scheme="None", force="None", tau=0.55, eps=0.0, store_every=50,
                           text="D3Q19 SRT-BGK 1024^3 strong-scaled, halo/interior overlap, on-line energy/enstrophy/Mach "
                                "reductions every 50 steps (BASELINE configs[4])"),
    "d3q19_bgk_512": dict(config=4, lattice="D3Q19", q=19, shape=(512, 512, 512), scaling="strong", collision="BGK",
                          equilibrium="TruncationMa3", scheme="None", force="None", tau=0.55, eps=0.0, store_every=50,
                          text="D3Q19 SRT-BGK 512^3 strong-scaled (reduced-size stand-in of BASELINE configs[4] that fits one GPU)"),
}


# Secondary measurements carried in the same JSON line under "also" (device-resident throughput only, measured after
# the headline numbers are final and under a watchdog, so that they can never change or lose the headline): the other
# BASELINE configs at this GPU count.  (workload, dtype, eps, stored mode, timed steps)
ALSO_SINGLE = [
    ("d3q19_bgk_256", "F32", None, 1, 100),
    ("d3q19_bgk_guo_256", "F64", None, 1, 100),
    ("d3q27_bgk_512", "F64", None, 1, 50),
    ("d3q27_elbm_512", "F64", 2e-2, 1, 20),
    ("d3q27_elbm_512", "F64", 1e-5, 1, 20),
    ("d3q27_elbm_512_resolved", "F64", None, 1, 20),   # no node leaves the shortcut: the HBM-bound end of the entropic kernel
    ("d2q9_elbm_shanchen_8192", "F64", 2e-2, 1, 50),
    ("d2q9_elbm_shanchen_8192", "F64", 1e-5, 1, 50),
    ("d2q9_elbm_edm_8192", "F32", 2e-2, 1, 50),
    ("d3q19_bgk_512", "F64", None, 1, 100),
]
ALSO_MULTI = [
    ("d3q19_bgk_256", "F64", None, 1, 100),    # weak scaling, 256^3 per GPU (the N > 1 headline of round 1)
    ("d3q19_bgk_512", "F64", None, 1, 100),    # stored fields + spectral enstrophy every 50 steps
    ("d3q27_elbm_512", "F64", 2e-2, 1, 20),
    ("d3q27_elbm_512", "F64", 1e-5, 1, 20),
    ("d2q9_elbm_shanchen_8192", "F64", 2e-2, 1, 50),
    ("d2q9_elbm_edm_8192", "F32", 2e-2, 1, 50),
    ("d3q19_bgk_1024", "F64", None, 1, 100),   # BASELINE configs[4] with whole fields and the spectral enstrophy (does not fit 2 GPUs)
]


def measured_peak():
    path = ROOT / "MEASURED_PEAKS.json"
    if path.is_file():
        try:
            return float(json.loads(path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_PEAK_GBS, "fallback (B200_PROFILING.md)"


def recorded_traffic():
    """dram bytes per launch of the fused kernel from the committed ncu capture, if there is one."""
    path = ROOT / "profiles" / "traffic.json"
    if path.is_file():
        try:
            return json.loads(path.read_text()).get("d3q19_bgk_f64_256_dram_bytes_per_launch")
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks and throttle reasons sampled DURING the timed region."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.device_index = device_index
        self.lines = []
        self.process = None
        self.thread = None

    def start(self):
        try:
            self.process = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.device_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.process = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.process.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, begin: float, end: float) -> dict:
        if self.process is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.process.terminate()
        try:
            self.process.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.process.kill()
        clocks, maxima, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [l for t, l in self.lines if begin - 0.05 <= t <= end + 0.15] or [l for _, l in self.lines]
        for row in rows:
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 9:
                continue
            try:
                clocks.append(float(parts[1]))
                maxima.append(float(parts[2]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[5:9]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(clocks) if clocks else None,
                "sm_max_mhz": max(maxima) if maxima else None,
                "reasons": sorted(reasons), "samples": len(clocks)}


def distributed_setup(gpus: int):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if gpus > 1 and world != gpus:
        raise SystemExit(f"--gpus {gpus} needs torchrun with --nproc-per-node {gpus} (WORLD_SIZE={world})")
    return rank, world, local_rank


# --------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation (oracle/_ref, compiled from /root/reference by
# oracle/refbuild.py in the build container) on the host cores, on a bounded x-slab sample of the workload
# --------------------------------------------------------------------------------------------------
def _host_cores() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def _reference_sample_text(refbuild, cfg, full: bool) -> str:
    planes = cfg.nx // cfg.nprocs
    edge = cfg.ny
    what = (f"the WHOLE {edge}^3 cube of BASELINE configs[1]" if full else
            f"bounded x-slab sample {cfg.nx}x{edge}x{edge} of the {edge}^3 cube")
    return (f"{what}: reference CPU build ({refbuild.TIMING_FLAGS}), {cfg.nprocs} forked MPI-shim ranks x {planes} x-planes of "
            f"{edge}x{edge} D3Q19 BGK FP64 nodes, timers of Algorithm::iterate (computation + communication)")


def run_reference(args) -> int:
    """The reference's own CPU implementation on the host cores, on this repository's arm's configuration.
    N = 1 (D3Q19 256^3): the whole cube, one x-slab per core; if K + W steps of it would not end within a few minutes on this
    host, the thin-slab sample (2 x-planes per rank).  N > 1 (D3Q19 1024^3 strong-scaled; 490 GB on a CPU): a bounded x-slab
    sample of the 1024 x 1024 cross-section, 2 planes per rank on at most 32 ranks."""
    rank, world, _ = distributed_setup(args.gpus)
    if rank != 0:
        return 0
    from oracle import refbuild
    cores = _host_cores()
    large = args.gpus > 1
    if large:
        full_cfg = None
        thin_cfg = refbuild.best_timing_config(min(cores, refbuild.LARGE_MAX_RANKS), refbuild.TIMING_PLANES_PER_RANK, refbuild.LARGE_EDGE)
        if thin_cfg is None:   # binaries of the large cross-section missing: the 256^2 sample rather than nothing
            large = False
    if not large:
        full_cfg = refbuild.best_timing_config(cores)
        thin_cfg = refbuild.best_timing_config(cores, refbuild.TIMING_PLANES_PER_RANK)
    if full_cfg is None and thin_cfg is None:
        print(json.dumps({"impl": "reference", "unavailable": "no prebuilt reference binary in oracle/_ref"}))
        return 0
    cfg, full = thin_cfg, False
    if full_cfg is not None:
        probe = refbuild.time_reference(full_cfg, 1, 1)
        if thin_cfg is None or probe["seconds"] * (args.steps + args.warmup) <= 150.0:
            cfg, full = full_cfg, True
    result = refbuild.time_reference(cfg, args.steps, args.warmup)
    sample = _reference_sample_text(refbuild, cfg, full)
    workload = ("D3Q19 SRT-BGK 1024^3 strong-scaled (BASELINE configs[4])" if large else
                "D3Q19 SRT-BGK periodic 256^3 FP64 per GPU (BASELINE configs[1])")
    line = {
        "impl": "reference", "metric": METRIC, "value": result["mlups"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": result["seconds"] / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong" if large else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{workload}; {sample}",
                   "lattice": LATTICE, "collision": "BGK", "host_cores": cores, "ranks": cfg.nprocs,
                   "global_length": [cfg.nx, cfg.ny, cfg.nz], "same_config_as_gpu_arm": full},
        "cpu_baseline": {"value": result["mlups"], "unit": UNIT, "cores": cfg.nprocs, "kind": "reference",
                         "sample": sample, "communication_share": result["communication_share"]},
        "e2e": {"value": result["mlups"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def cpu_baseline_leg() -> dict:
    """The reference CPU build timed on this box's host cores: about 12 s on the whole 256^3 cube (one x-slab per core), and
    the thin-slab sample of round 1 (2 x-planes per rank, where the reference's halo exchange and boundary copies weigh as
    much as its node update) beside it."""
    from oracle import refbuild
    cores = _host_cores()

    def timed(cfg, budget):
        probe = refbuild.time_reference(cfg, 2, 1)
        steps = int(max(3, min(200, budget / max(probe["seconds"] / 2, 1e-3))))
        return refbuild.time_reference(cfg, steps, 1), steps

    full_cfg = refbuild.best_timing_config(cores)
    thin_cfg = refbuild.best_timing_config(cores, refbuild.TIMING_PLANES_PER_RANK)
    if full_cfg is None and thin_cfg is None:
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "no prebuilt reference binary"}
    out = None
    if full_cfg is not None:
        result, steps = timed(full_cfg, 12.0)
        out = {"value": result["mlups"], "unit": UNIT, "cores": full_cfg.nprocs, "kind": "reference",
               "sample": f"{steps} steps of " + _reference_sample_text(refbuild, full_cfg, True) + f", on {cores} host cores",
               "communication_share": result["communication_share"]}
    if thin_cfg is not None:
        result, steps = timed(thin_cfg, 5.0)
        thin = {"value": result["mlups"], "unit": UNIT, "cores": thin_cfg.nprocs,
                "sample": f"{steps} steps of " + _reference_sample_text(refbuild, thin_cfg, False),
                "communication_share": result["communication_share"]}
        if out is None:
            out = {**thin, "kind": "reference"}
        else:
            out["thin_slab_sample"] = thin
    return out


# --------------------------------------------------------------------------------------------------
# secondary workloads ("also")
# --------------------------------------------------------------------------------------------------
def device_bytes_needed(q_count, plane_nodes, lx, element, entropic, dim, stored_mode) -> int:
    """Device memory of a workload: two population buffers with their x-halo planes, the alpha field of the entropic
    collisions and, when whole fields are stored, density / velocity / force plus the three spectral work arrays."""
    nodes_local = plane_nodes * lx
    need = 2 * q_count * plane_nodes * (lx + 2) * element + (nodes_local * element if entropic else 0)
    if stored_mode == 1:
        need += (1 + 2 * dim) * nodes_local * element + 3 * dim * nodes_local * 8
    return need


def fp64_peak():
    """DFMA lane-operations per second measured by tools/fp64_peak.cu on this pool's B200 (profiles/FP64_PEAK.json;
    MEASURED_PEAKS.json has no FP64 figure), else the nominal 64 DFMA per clock and SM at 1965 MHz."""
    path = ROOT / "profiles" / "FP64_PEAK.json"
    if path.is_file():
        try:
            return float(json.loads(path.read_text())["dfma_lane_ops_per_s"]), "measured (profiles/FP64_PEAK.json, tools/fp64_peak.cu)"
        except Exception:
            pass
    return 148 * 64 * 1.965e9, "nominal (64 DFMA / clock / SM x 148 SMs x 1965 MHz)"


# FP64-pipe instructions per node of the shipped entropic kernels, counted from their SASS by scripts/fp64_counts.py
# (profiles/r02_sass_fp64_counts.md): `base` = every node (pull, moments, equilibrium, screens, collide), `solve` = once per node
# that leaves the shortcut (alphaMax where the solving thread forms it, the hoisted sum of f ln f), `evaluation` = per evaluation
# of (F, F') of the Newton solve.  The FP64 side of the entropic roofline; one warp instruction = 32 lane operations = 32 nodes.
FP64_OPS = {
    # workload: (base, solve, evaluation)
    "d3q27_elbm_512": (780, 270, 309),
    "d3q27_elbm_512_resolved": (780, 270, 309),
    "d2q9_elbm_shanchen_8192": (211, 136, 123),
    "d2q9_elbm_edm_8192": (290, 136, 123),
}


def fp64_fraction(workload, kernel_nodes, kernel_ms, newton) -> dict | None:
    counts = FP64_OPS.get(workload)
    if not counts or not kernel_ms:
        return None
    base, solve, evaluation = counts
    solved = (newton or {}).get("solved_node_fraction", 0.0)
    per_solved = (newton or {}).get("evaluations_per_solved_node", 0.0)
    per_node = base + solved * (solve + per_solved * evaluation)
    peak, source = fp64_peak()
    achieved = per_node * kernel_nodes / (kernel_ms * 1e-3)
    return {"fp64_ops_per_node": per_node, "fp64_achieved_lane_ops_per_s": achieved, "fp64_peak_lane_ops_per_s": peak,
            "fp64_frac": achieved / peak, "fp64_peak_source": source}


def bind_near_gpu(device_index: int):
    """Restrict this process to the host cores of the GPU's NUMA node, so that the pinned staging buffer of the end-to-end
    leg is allocated (first touch) in the memory next to the GPU's PCIe root.  Returns the previous affinity (or None)."""
    try:
        import torch
        props = torch.cuda.get_device_properties(device_index)
        address = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        node = int((Path("/sys/bus/pci/devices") / address / "numa_node").read_text())
        if node < 0:
            return None
        cpus = set()
        for part in (Path("/sys/devices/system/node") / f"node{node}" / "cpulist").read_text().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        before = os.sched_getaffinity(0)
        wanted = cpus & before
        if not wanted:
            return None
        os.sched_setaffinity(0, wanted)
        return before
    except Exception:  # noqa: BLE001 -- placement is best effort
        return None


def host_memory_available() -> int:
    try:
        for row in Path("/proc/meminfo").read_text().splitlines():
            if row.startswith("MemAvailable:"):
                return int(row.split()[1]) * 1024
    except Exception:  # noqa: BLE001
        pass
    return 0


def summarize_also(also) -> list:
    """One short string per secondary workload: the line's LAST key, so that it survives a truncated log tail."""
    rows = []
    for entry in also:
        try:
            if "value" not in entry:
                rows.append(f"{entry.get('name')}: {entry.get('skipped') or entry.get('error') or '?'}"[:90])
                continue
            length = entry.get("global_length") or ["?"] * 3
            text = f"{entry.get('name')} {entry.get('dtype')} {length[0]}x{length[1]}x{length[2]}"
            if entry.get("eps"):
                text += f" eps={entry['eps']:g}"
            if entry.get("stored_mode"):
                text += f" store{entry['stored_mode']}"
            text += f": {entry['value']:.0f} MLUPS"
            if entry.get("roofline_frac") is not None:
                text += f" hbm={entry['roofline_frac']:.3f}"
            if entry.get("fp64_frac") is not None:
                text += f" fp64={entry['fp64_frac']:.2f}"
            if entry.get("stored_step_ms") is not None and entry.get("ms_per_step") is not None:
                text += f" stored_ms={entry['stored_step_ms']:.2f}/{entry['ms_per_step']:.2f}"
            rows.append(text)
        except Exception as error:  # noqa: BLE001 -- a summary must never cost the line
            rows.append(f"{entry.get('name') if isinstance(entry, dict) else entry}: summary failed ({error})"[:90])
    return rows


def newton_work(algorithm, iteration, nodes_local) -> dict:
    """One extra (untimed) step with the kernel's Newton counters on: which share of this rank's nodes took the solve
    (solveAlpha, Collision.h:328-349) and how many evaluations of (F, F') a solved node needed -- the FP64 side of the
    entropic roofline, at the END of the timed region."""
    algorithm.newton_statistics(start=True)
    algorithm.run(iteration, 1, 0)
    solved, evaluations = algorithm.newton_statistics()
    return {"solved_node_fraction": solved / nodes_local, "evaluations_per_solved_node": evaluations / solved if solved else 0.0}


def measure_also(entry, args, rank, world, local_rank, barrier, max_over_ranks, min_over_ranks, peak) -> dict:
    """Device-resident throughput of one named workload at this GPU count: fresh context, synthetic field made on
    the device, 5 warm-up steps, `steps` timed steps between barriers (CUDA events, max over ranks)."""
    import torch
    from metalbm_b200.algorithm import Algorithm, Communication
    from metalbm_b200.capi import make_config

    name, dtype, eps, stored_mode, steps = entry
    work = dict(WORKLOADS[name])
    if eps is not None:
        work["eps"] = eps
    element = 8 if dtype == "F64" else 4
    q_count = work["q"]
    entropic = work["collision"] != "BGK"
    shape = (work["shape"][0] * world,) + tuple(work["shape"][1:]) if work["scaling"] == "weak" else tuple(work["shape"])
    result = {"name": name, "dtype": dtype.lower(), "eps": work["eps"], "scaling": work["scaling"], "n_gpus": world,
              "global_length": list(shape), "store_every": int(work["store_every"]),
              "stored_mode": stored_mode if work["store_every"] else None}
    if shape[0] % world:
        result["skipped"] = "number of GPUs does not divide the x extent"
        return result
    nodes_global = shape[0] * shape[1] * shape[2]
    nodes_local = nodes_global // world
    lx = shape[0] // world
    dim = 3 if shape[2] > 1 else 2
    # device memory this workload needs (populations with halo planes, alpha, and on fully stored steps the fields
    # plus the spectra of the enstrophy), agreed over the ranks before anybody allocates
    need = device_bytes_needed(q_count, nodes_local // lx, lx, element, entropic, dim, stored_mode if work["store_every"] else 0)
    free = torch.cuda.mem_get_info()[0]
    fits = min_over_ranks(1.0 if need + (3 << 30) < free else 0.0)
    result["device_bytes_needed"] = need
    if not fits:
        result["skipped"] = f"needs {need / 1e9:.1f} GB of the {free / 1e9:.1f} GB free per GPU"
        return result
    bytes_per_node = 2 * q_count * element + (2 * element if entropic else 0)
    cfg = make_config(lattice=work["lattice"], shape=shape, collision=work["collision"], equilibrium=work["equilibrium"],
                      forcing_scheme=work["scheme"], force=work["force"], tau=work["tau"], dtype=dtype,
                      amplitude=(1e-5, 1e-5, 1e-5), wavelength=(32.0, 32.0, 32.0),
                      overlap=args.overlap, rank=rank, nranks=world, device=local_rank, variant=args.variant)
    algorithm = Algorithm(cfg, communication=Communication(rank, world), host_distribution=False, host_fields=False,
                          peer_halos=(args.halo == "peer" and args.overlap == "On") if world > 1 else False)
    try:
        algorithm.init_synthetic(work.get("flow", 0.05), work.get("flow", 0.05))
        if work["eps"]:
            algorithm.perturb(work["eps"])
        store_every = int(work["store_every"])
        warmup = 5
        if store_every:
            algorithm.run(0, 1, 1, stored_mode=stored_mode)      # first stored step: allocations, transform plans
        algorithm.run(1, warmup, store_every, stored_mode=stored_mode)
        algorithm.kernel_time()
        barrier()
        algorithm.mark(0)
        algorithm.run(warmup + 1, steps, store_every, sync=False, stored_mode=stored_mode)
        algorithm.mark(1)
        algorithm.synchronize()
        barrier()
        device_ms = max_over_ranks(algorithm.elapsed_ms(0, 1))
        kernel_ms, kernel_launches = algorithm.kernel_time()
        stored_ms = None
        if store_every:
            barrier()
            algorithm.mark(2)
            algorithm.run(store_every, 1, store_every, sync=False, stored_mode=stored_mode)
            algorithm.mark(3)
            algorithm.synchronize()
            stored_ms = max_over_ranks(algorithm.elapsed_ms(2, 3))
        # what the entropic collision was doing at the end of the timed region: share of nodes off the alpha = 2 shortcut
        newton_fraction, alpha_min, alpha_max = algorithm.alpha_statistics() if entropic else (None, None, None)
        newton = newton_work(algorithm, warmup + steps + 1, nodes_local) if entropic else None
        # liveness of the state that was timed: one more step reducing energy / mass / Mach only
        algorithm.run(0, 1, 1, stored_mode=2)
        observables = algorithm.observables()
        overlapped = world > 1 and args.overlap == "On" and lx >= 3
        kernel_nodes = nodes_local // lx * (lx - 2) if overlapped else nodes_local
        achieved = bytes_per_node * kernel_nodes / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else None
        result.update({
            "value": nodes_global * steps / (device_ms * 1e-3) / 1e6, "unit": UNIT, "steps": steps, "warmup": warmup,
            "ms_per_step": device_ms / steps, "stored_step_ms": stored_ms,
            "kernel_ms": kernel_ms, "kernel_launches_timed": kernel_launches,
            "roofline_achieved_GBps": achieved, "roofline_frac": achieved / peak if achieved else None,
            "roofline_mlups_per_gpu": peak * 1e9 / bytes_per_node / 1e6,
            "halo": ("peer" if algorithm.peer_halos else "nccl") if world > 1 else "none",
            "energy": float(observables[0]), "mach": float(observables[2]),
            "mass_per_node": float(observables[3]) / nodes_global,
        })
        if entropic:
            result.update({"alpha_off_shortcut_fraction_at_end": newton_fraction, "alpha_min": alpha_min, "alpha_max": alpha_max,
                           "newton": newton})
            fp64 = fp64_fraction(name, kernel_nodes, kernel_ms, newton)
            if fp64:
                result.update(fp64)
    finally:
        algorithm.close()
    return result


def run_secondary(entries, measure, line, rank, world, max_over_ranks, sum_over_ranks, timeout, stream=None) -> None:
    """Run the secondary workloads and print THE JSON line (rank 0) exactly once, whatever happens to them.

    `line` (rank 0: the finished headline dict, other ranks: None) is printed with the results under "also":
      * every entry is measured inside try/except, its failure is an {"error": ...} entry;
      * with several ranks, a SUM all-reduce after every entry returns only when all ranks are done with it, so the ranks
        stay in step after symmetric failures (not enough memory, a missing library);
      * a rank that never comes back (a neighbour died inside a collective) is ended by a watchdog thread after `timeout`
        seconds: it prints the line with what was measured so far and leaves with exit code 0;
      * a rank on which the bench's own collectives fail publishes the line at once and leaves when the watchdogs of the
        other ranks fire, so that no process disappears under a running neighbour.
    Pure host logic (tests/test_bench_host.py drives it with fake workloads)."""
    stream = stream or sys.stdout
    also = []
    printed = threading.Lock()

    def emit(note=None):
        if not printed.acquire(blocking=False):
            return
        if rank == 0:
            if also or note:
                line["also"] = list(also)
            if note:
                line["also_note"] = note
            if also:
                line["also_summary"] = summarize_also(also)
            stream.write(json.dumps(line) + "\n")
            stream.flush()

    if not entries:
        emit()
        return

    def bail():
        emit(f"secondary workloads cut off by the {timeout} s watchdog")
        os._exit(0)

    watchdog = threading.Timer(timeout, bail)
    watchdog.daemon = True
    watchdog.start()
    started = time.time()
    in_step = True   # the ranks still take the same decisions
    try:
        for entry in entries:
            # every rank takes the same decision: the elapsed time is agreed first
            if max_over_ranks(time.time() - started) > timeout - min(60.0, 0.2 * timeout):
                also.append({"name": entry[0], "skipped": "time budget of the secondary workloads used up"})
                continue
            ok = 1.0
            try:
                also.append(measure(entry))
            except Exception as error:  # noqa: BLE001 -- reported in the line, the headline stands
                ok = 0.0
                also.append({"name": entry[0], "dtype": str(entry[1]).lower(), "eps": entry[2], "n_gpus": world,
                             "error": str(error)[:300]})
            if world > 1:
                # returns once every rank is done with this entry, whatever its outcome: the ranks are in step again
                ranks_ok = sum_over_ranks(ok)
                if ranks_ok != world:
                    also[-1]["ranks_ok"] = int(ranks_ok)
    except Exception as error:  # noqa: BLE001 -- a collective of the bench itself failed (e.g. a poisoned CUDA context)
        in_step = False
        also.append({"error": f"secondary workloads stopped: {str(error)[:300]}"})
    if in_step:
        watchdog.cancel()
        emit()
        return
    emit("secondary workloads stopped after an error on this rank")
    if world > 1:
        time.sleep(max(0.0, started + timeout - time.time()) + 5.0)   # the watchdog ends this process
    os._exit(0)


# --------------------------------------------------------------------------------------------------
# this repository's arm
# --------------------------------------------------------------------------------------------------
def run_ours(args) -> int:
    import numpy as np
    import torch
    import torch.distributed as dist

    from metalbm_b200.algorithm import Algorithm, Communication
    from metalbm_b200.capi import check, make_config

    rank, world, local_rank = distributed_setup(args.gpus)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(value: float) -> float:
        if world == 1:
            return value
        tensor = torch.tensor([value], dtype=torch.float64, device="cuda")
        dist.all_reduce(tensor, op=dist.ReduceOp.MAX)
        return float(tensor.item())

    def min_over_ranks(value: float) -> float:
        return -max_over_ranks(-value)

    def sum_over_ranks(value: float) -> float:
        if world == 1:
            return value
        tensor = torch.tensor([value], dtype=torch.float64, device="cuda")
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
        return float(tensor.item())

    # N = 1: BASELINE configs[1] (256^3, the roofline calibration).  N > 1: BASELINE configs[4], D3Q19 1024^3 strong-scaled
    # -- the configuration the north star's parallel-efficiency target is quoted on; the driver's efficiency
    # v_N / (N v_1) is then SURVEY 8d's definition (1024^3 itself cannot run on one GPU).
    workload_name = args.workload or ("d3q19_bgk_256" if world == 1 else "d3q19_bgk_1024")
    default_headline = (workload_name == ("d3q19_bgk_256" if world == 1 else "d3q19_bgk_1024") and args.dtype == "f64"
                        and args.edge == EDGE and not getattr(args, "shape", None))
    work = dict(WORKLOADS[workload_name])
    if args.edge != EDGE and workload_name == "d3q19_bgk_256":
        work["shape"] = (args.edge,) * 3
        work["text"] = work["text"].replace("256^3", f"{args.edge}^3")
    if getattr(args, "shape", None):
        work["shape"] = tuple(int(n) for n in args.shape.split(","))
        work["scaling"] = "strong"
        work["text"] += f" [shape overridden: {args.shape}]"
    if args.eps is not None:
        work["eps"] = args.eps
    if args.store_every is not None:
        work["store_every"] = args.store_every
    dtype = args.dtype.upper()
    element = 8 if dtype == "F64" else 4
    q_count = work["q"]
    entropic = work["collision"] != "BGK"
    bytes_per_node = 2 * q_count * element + (2 * element if entropic else 0)
    if work["scaling"] == "weak":
        shape = (work["shape"][0] * world,) + tuple(work["shape"][1:])
    else:
        shape = tuple(work["shape"])
    if shape[0] % world:
        raise SystemExit(f"{world} GPUs do not divide the x extent {shape[0]}")
    nodes_global = shape[0] * shape[1] * shape[2]
    nodes_local = nodes_global // world
    lx = shape[0] // world
    dim = 3 if shape[2] > 1 else 2
    store_every = int(work["store_every"])
    # the slab has to fit: 1024^3 on 2 GPUs takes 164 of a B200's ~190 GB free.  Should a box offer less, the headline falls back
    # to the 512^3 strong-scaled stand-in rather than to no line at all, and says so.
    bare = device_bytes_needed(q_count, nodes_local // lx, lx, element, entropic, dim, 2) + (6 << 30)
    if default_headline and world > 1 and not min_over_ranks(1.0 if bare < torch.cuda.mem_get_info()[0] else 0.0):
        fallback = dict(WORKLOADS["d3q19_bgk_512"])
        fallback["text"] += f" [FALLBACK: the 1024^3 slab needs {bare / 1e9:.0f} GB per GPU, more than this box has free]"
        work, workload_name, shape = fallback, "d3q19_bgk_512", tuple(fallback["shape"])
        nodes_global = shape[0] * shape[1] * shape[2]
        nodes_local = nodes_global // world
        lx = shape[0] // world
        store_every = int(work["store_every"])
    # stored steps: whole fields + energy / spectral enstrophy / Mach (mode 1) where the slab leaves room for the field
    # arrays, else the energy / mass / Mach reductions alone (mode 2: no field arrays; 1024^3 on 2 GPUs fills them)
    stored_mode = 0
    if store_every:
        free = torch.cuda.mem_get_info()[0]
        full = device_bytes_needed(q_count, nodes_local // lx, lx, element, entropic, dim, 1) + (4 << 30) < free
        stored_mode = getattr(args, "stored_mode", 0) or (1 if min_over_ranks(1.0 if full else 0.0) else 2)
    cfg = make_config(lattice=work["lattice"], shape=shape, collision=work["collision"], equilibrium=work["equilibrium"],
                      forcing_scheme=work["scheme"], force=work["force"], tau=work["tau"], dtype=dtype,
                      amplitude=(1e-5, 1e-5, 1e-5), wavelength=(32.0, 32.0, 32.0),
                      overlap=args.overlap, rank=rank, nranks=world, device=local_rank, variant=args.variant)
    algorithm = Algorithm(cfg, communication=Communication(rank, world), host_distribution=False, host_fields=False,
                          peer_halos=(args.halo == "peer" and args.overlap == "On") if world > 1 else False)
    domain = algorithm.domain

    # synthetic initial field of the named grid size, made on the device (SURVEY 8d "Init B": Taylor-Green-like velocity,
    # density ripple, f = feq(rho, u)), then (entropic workloads) a multiplicative perturbation of relative size eps
    algorithm.init_synthetic(work.get("flow", 0.05), work.get("flow", 0.05))
    if work["eps"]:
        algorithm.perturb(work["eps"])

    # ---- device-resident throughput: W warm-up steps, then exactly K timed steps -------------------------
    if store_every:
        # the first stored step allocates the field arrays and plans the transforms of the spectral enstrophy: warm-up.
        # Should that fail on any rank (no room for the transform's work area), every rank falls back to the reductions alone.
        failure = None
        try:
            algorithm.run(0, 1, 1, stored_mode=stored_mode)
        except Exception as error:  # noqa: BLE001 -- agreed over the ranks below
            failure = str(error)[:200]
        if min_over_ranks(0.0 if failure else 1.0) <= 0:
            if stored_mode == 2:
                raise RuntimeError(f"the first stored step failed: {failure or 'on another rank'}")
            stored_mode = 2
            work["text"] += f" [stored steps fell back to the reductions alone: {failure or 'failure on another rank'}]"
            algorithm.run(0, 1, 1, stored_mode=stored_mode)
    algorithm.run(1, args.warmup, store_every, stored_mode=stored_mode or 1)
    algorithm.kernel_time()                      # switches the per-launch CUDA event pairs on
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches_before = algorithm.launch_count()
    barrier()
    begin = time.time()
    algorithm.mark(0)
    algorithm.run(args.warmup + 1, args.steps, store_every, sync=False, stored_mode=stored_mode or 1)
    algorithm.mark(1)
    algorithm.synchronize()
    barrier()
    end = time.time()
    device_ms = max_over_ranks(algorithm.elapsed_ms(0, 1))
    launches = algorithm.launch_count() - launches_before
    kernel_ms, kernel_launches = algorithm.kernel_time()
    clocks = sampler.stop(begin, end) if rank == 0 else None
    value = nodes_global * args.steps / (device_ms * 1e-3) / 1e6
    stored_in_region = sum(1 for i in range(args.warmup + 1, args.warmup + args.steps + 1) if store_every and i % store_every == 0)
    entropic_state = None
    if entropic:
        off_shortcut, alpha_min, alpha_max = algorithm.alpha_statistics()
        entropic_state = {"alpha_off_shortcut_fraction_at_end": off_shortcut, "alpha_min": alpha_min, "alpha_max": alpha_max,
                          "newton": newton_work(algorithm, args.warmup + args.steps + 1, nodes_local)}

    # ---- cost of one stored step (fields + energy / spectral enstrophy / Mach reductions), timed on its own -------------
    stored_ms = stored_with_analysis_ms = None
    if store_every:
        barrier()
        algorithm.mark(2)
        algorithm.run(store_every, 1, store_every, sync=False, stored_mode=stored_mode)
        algorithm.mark(3)
        algorithm.synchronize()
        stored_ms = max_over_ranks(algorithm.elapsed_ms(2, 3))
        if stored_mode == 1:
            # the spectral analysis of that step ran on its own stream: reading the observables joins it, so 2 -> 4 is the
            # latency of the stored step INCLUDING its analysis (what a caller that reads the enstrophy at once waits for)
            barrier()
            algorithm.mark(2)
            algorithm.run(store_every, 1, store_every, sync=False, stored_mode=stored_mode)
            algorithm.observables()
            algorithm.mark(4)
            algorithm.synchronize()
            stored_with_analysis_ms = max_over_ranks(algorithm.elapsed_ms(2, 4))

    peak, peak_source = measured_peak()
    # the dominant kernel is the bulk launch of the fused step: all local planes at N = 1, all but the two boundary
    # planes when the halo exchange is overlapped
    overlapped = world > 1 and args.overlap == "On" and lx >= 3
    kernel_nodes = nodes_local // lx * (lx - 2) if overlapped else nodes_local
    achieved = bytes_per_node * kernel_nodes / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else None
    kernel_name = (f"fusedStepKernel<{work['lattice']},{work['collision']},{work['equilibrium']},{work['scheme']},"
                   f"{'double' if element == 8 else 'float'}>")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak if achieved else None,
                # ncu's dram bytes per launch exist for exactly one kernel and shape (profiles/traffic.json says which)
                "traffic": recorded_traffic() if workload_name == "d3q19_bgk_256" and dtype == "F64" and world == 1
                and args.edge == EDGE and not getattr(args, "shape", None) else None,
                "peak_source": peak_source, "kernel": kernel_name,
                "kernel_ms": kernel_ms, "kernel_launches_timed": kernel_launches,
                "kernel_planes_per_launch": lx - 2 if overlapped else lx,
                "algorithmic_bytes_per_node": bytes_per_node,
                "algorithmic_bytes_per_launch": bytes_per_node * kernel_nodes,
                "roofline_mlups_per_gpu": peak * 1e9 / bytes_per_node / 1e6}
    if entropic_state is not None:
        fp64 = fp64_fraction(workload_name, kernel_nodes, kernel_ms, entropic_state["newton"])
        if fp64:
            roofline.update(fp64)

    algorithm_peer = algorithm.peer_halos

    line = None
    if rank == 0:
        buffer_gb = q_count * nodes_local * element / 1e9
        metric = METRIC if work["lattice"] == "D3Q19" and dtype == "F64" else \
            f"MLUPS ({work['lattice']}, {'FP64' if element == 8 else 'FP32'})"
        stored_text = {0: None, 1: "whole fields + energy / spectral enstrophy / Mach", 2: "energy / mass / Mach reductions, no field arrays"}
        line = {
            "metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": device_ms / args.steps, "higher_is_better": True, "scaling": work["scaling"], "vs_baseline": None,
            "dtype": "f64" if element == 8 else ("f32 storage, f64 arithmetic" if entropic else "f32"), "data": "synthetic",
            "config": {"workload": f"{work['text']}; global {shape[0]}x{shape[1]}x{shape[2]}", "name": workload_name,
                       "lattice": work["lattice"], "collision": work["collision"], "equilibrium": work["equilibrium"],
                       "forcing": f"{work['scheme']}/{work['force']}", "tau": work["tau"], "perturbation_eps": work["eps"],
                       "store_every": store_every, "stored_mode": stored_text[stored_mode], "stored_step_ms": stored_ms,
                       "stored_step_with_analysis_ms": stored_with_analysis_ms,
                       "stored_steps_in_timed_region": stored_in_region,
                       "global_length": list(shape), "parallelism": f"x-slab x{world}",
                       "overlap": args.overlap,
                       "halo": ("direct peer stores over NVLink from the boundary kernel" if algorithm_peer else
                                "NCCL send/recv") if world > 1 else "none (single rank: periodic wrap in the kernel)",
                       "l2": f"inputs (2 x {buffer_gb:.2f} GB population buffers per GPU) larger than the 126 MB L2; "
                             "no flush between steps"},
            "clocks": clocks, "e2e": None, "gpu_launches": launches, "roofline": roofline,
        }
        if store_every and stored_ms is not None:
            # the configuration's cadence spelled out: K plain steps were timed; one step in store_every is a stored one
            plain = device_ms / args.steps if not stored_in_region else None
            if plain is not None:
                line["config"]["mlups_at_cadence"] = nodes_global * store_every / ((plain * (store_every - 1) + stored_ms) * 1e-3) / 1e6
        if entropic_state is not None:
            line["entropic"] = entropic_state

    # ---- end to end through the C-ABI with host buffers ---------------------------------------------------
    # After the headline is final and under a watchdog: should the leg stop answering (a rank lost inside a collective, a host
    # allocation that never returns) the line is printed without it, by every rank's timer, instead of never.
    def end_to_end() -> dict:
        def agree(ok: bool, what: str) -> None:
            """every rank learns whether ALL ranks got through a step that can fail on one of them alone"""
            if min_over_ranks(1.0 if ok else 0.0) <= 0:
                raise RuntimeError(what)

        affinity = None
        try:
            # the host side of Distribution<T, GPU> (Distribution.h:15-43): the local padded SoA array, in pinned memory
            # next to this GPU's PCIe root
            host_bytes = q_count * domain.number_elements * element
            available = host_memory_available()
            # pinned where the ranks' buffers take less than 45 % of the host's free memory, pageable up to 80 % (1024^3 on 2
            # GPUs: 2 x 82 GB), else no end-to-end leg
            pin = min_over_ranks(1.0 if (available == 0 or host_bytes * world < 0.45 * available) else 0.0) > 0
            fits = min_over_ranks(1.0 if (available == 0 or host_bytes * world < 0.80 * available) else 0.0) > 0
            if not fits:
                raise RuntimeError(f"host staging buffers of {host_bytes * world / 1e9:.0f} GB do not fit {available / 1e9:.0f} GB of host memory")
            affinity = bind_near_gpu(local_rank)
            failure = None
            try:
                if pin:
                    pinned = torch.empty((q_count,) + domain.padded_length, dtype=torch.float64 if element == 8 else torch.float32,
                                         pin_memory=True)
                    algorithm.distribution.array = pinned.numpy()
                else:
                    algorithm.distribution.array = np.empty((q_count,) + domain.padded_length, dtype=domain.dtype)
                algorithm.pack()                     # current state -> host array (first-touches the pages)
            except Exception as error:  # noqa: BLE001 -- reported by every rank below
                failure = str(error)[:200]
            agree(failure is None, f"host buffer or first download failed on a rank ({failure or 'another rank'})")
            # warm-up of everything the timed region uses for the first time: the observables' partial sums and their
            # all-reduce (NCCL sets its channels up on the first collective of a communicator), the SM clocks
            for iteration in range(1, 1 + max(args.warmup, 20)):
                check(algorithm._lib.mlbm_step(algorithm._ctx, iteration, 2))
                algorithm.observables()
            barrier()
            t0 = time.perf_counter()
            algorithm.unpack()                       # H2D of the whole SoA distribution from pinned host memory
            t1 = time.perf_counter()
            energy = 0.0
            for iteration in range(1, args.steps + 1):
                algorithm._lib.mlbm_step(algorithm._ctx, iteration, 2)   # isStored = observables only
                energy = algorithm.observables()[0]  # D2H read of the step's scalar results
            t2 = time.perf_counter()
            algorithm.pack()                         # D2H of the whole distribution
            t3 = time.perf_counter()
            barrier()
            e2e_seconds = max_over_ranks(time.perf_counter() - t0)
            e2e_value = nodes_global * args.steps / e2e_seconds / 1e6
            distribution_bytes = q_count * nodes_local * element
            return {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": distribution_bytes / args.steps,
                    "d2h_bytes_per_step": distribution_bytes / args.steps + 32,
                    "note": "per rank; timed region: unpack (H2D of all populations from pinned host memory) + K synchronous "
                            "mlbm_step calls each followed by a D2H read of the all-reduced observables + pack (D2H of all "
                            "populations); population bytes amortised over K; max over ranks", "last_energy": energy,
                    "unpack_ms": max_over_ranks((t1 - t0) * 1e3), "steps_ms": max_over_ranks((t2 - t1) * 1e3),
                    "pack_ms": max_over_ranks((t3 - t2) * 1e3),
                    "h2d_GBps": distribution_bytes / (t1 - t0) / 1e9, "d2h_GBps": distribution_bytes / (t3 - t2) / 1e9,
                    "host_buffer": "pinned" if pin else "pageable (pinning it would take more than 45 % of the host's free memory)",
                    "host_buffer_numa_bound": affinity is not None}
        except Exception as error:  # noqa: BLE001 -- the device-resident headline above stands; the line says what happened
            return {"value": None, "unit": UNIT, "error": str(error)[:300]}
        finally:
            if affinity is not None:
                os.sched_setaffinity(0, affinity)

    if not args.no_e2e:
        timeout = float(getattr(args, "e2e_timeout", 300))

        def give_up():
            if rank == 0:
                line["e2e"] = {"value": None, "unit": UNIT, "error": f"the end-to-end leg did not finish within {timeout:.0f} s"}
                sys.stdout.write(json.dumps(line) + "\n")
                sys.stdout.flush()
            os._exit(0)

        watchdog = threading.Timer(timeout, give_up)
        watchdog.daemon = True
        watchdog.start()
        e2e = end_to_end()
        watchdog.cancel()
        if rank == 0:
            line["e2e"] = e2e

    cpu = cpu_baseline_leg() if (rank == 0 and world == 1 and not args.no_cpu_baseline
                                 and workload_name == "d3q19_bgk_256") else None
    if rank == 0 and cpu is not None:
        line["cpu_baseline"] = cpu
    algorithm.close()

    # ---- the other BASELINE configs at this GPU count.  The headline line above is final; whatever happens below
    # (an exception, a rank that stops answering) it is printed, by the watchdog if need be.
    run_also = args.also == "on" or (args.also == "auto" and default_headline)
    entries = (ALSO_SINGLE if world == 1 else ALSO_MULTI) if run_also else []

    def measure(entry):
        return measure_also(entry, args, rank, world, local_rank, barrier, max_over_ranks, min_over_ranks, peak)

    run_secondary(entries, measure, line, rank, world, max_over_ranks, sum_over_ranks, args.also_timeout)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main() -> int:
    parser = argparse.ArgumentParser()
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--steps", type=int, default=200)
    parser.add_argument("--warmup", type=int, default=10)
    parser.add_argument("--impl", default="ours", choices=["ours", "reference"])
    parser.add_argument("--edge", type=int, default=EDGE, help="edge of the per-GPU cube (default: the BASELINE 256)")
    parser.add_argument("--variant", type=int, default=0)
    parser.add_argument("--shape", default=None, help="global X,Y,Z overriding the workload's grid (experiments: slab shapes)")
    parser.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                        help="default: d3q19_bgk_256 on one GPU (BASELINE configs[1]), d3q19_bgk_1024 strong-scaled on several (configs[4])")
    parser.add_argument("--stored-mode", type=int, default=0, choices=[0, 1, 2],
                        help="stored steps: 1 = whole fields + spectral enstrophy, 2 = energy / mass / Mach only; 0 = 1 where it fits")
    parser.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    parser.add_argument("--overlap", default="On", choices=["On", "Off"])
    parser.add_argument("--halo", default="peer", choices=["peer", "nccl"],
                        help="N > 1 with overlap On: boundary kernel stores into the neighbours' halo planes (peer) or NCCL send/recv")
    parser.add_argument("--eps", type=float, default=None, help="override the workload's initial perturbation")
    parser.add_argument("--store-every", type=int, default=None, help="override the workload's observable cadence")
    parser.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg (large workloads)")
    parser.add_argument("--no-cpu-baseline", action="store_true")
    parser.add_argument("--also", default="auto", choices=["auto", "on", "off"],
                        help="secondary workloads (the other BASELINE configs at this GPU count) under \"also\" in the "
                             "JSON line; auto = with the default headline workload only")
    parser.add_argument("--e2e-timeout", type=int, default=300, help="watchdog of the host-buffer end-to-end leg, seconds")
    parser.add_argument("--also-timeout", type=int, default=300, help="watchdog of the secondary workloads, seconds")
    args = parser.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
