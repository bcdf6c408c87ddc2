"""ctypes binding of the C-ABI declared in ``include/metalbm_b200.h``.

This is the same boundary the C++ template shim (``include/metaLBM_b200/``) calls; Python uses
it for the parity tests and the benchmark driver.  There is no fallback: if the CUDA library
is missing, loading fails loudly.
"""
from __future__ import annotations

import ctypes
import enum
import os
from pathlib import Path

PACKAGE_DIR = Path(__file__).resolve().parent
# MLBM_VARIANT selects an experimental build (metalbm_b200/build.py); the product library otherwise
LIBRARY_PATH = PACKAGE_DIR / ("libmetalbm_b200" + ("_" + os.environ["MLBM_VARIANT"] if os.environ.get("MLBM_VARIANT") else "") + ".so")
ABI_VERSION = 2
PEER_HANDLE_BYTES = 256


class Lattice(enum.IntEnum):
    D2Q5 = 0
    D2Q9 = 1
    D3Q15 = 2
    D3Q19 = 3
    D3Q27 = 4
    D2Q13 = 5   # multi-speed lattices (halo 2-3): one GPU
    D2Q17 = 6
    D2Q21 = 7
    D3Q33 = 8


class Collision(enum.IntEnum):
    BGK = 0
    ELBM = 1
    ForcedNR_ELBM = 2
    Approached_ELBM = 3
    Malaspinas_ELBM = 4
    Essentially1_ELBM = 5
    Essentially2_ELBM = 6
    ForcedBNR_ELBM = 7
    ForcedNR_ELBM_Forcing = 8


class Equilibrium(enum.IntEnum):
    TruncationMa3 = 0
    Exact = 1


class ForcingScheme(enum.IntEnum):
    None_ = 0
    Guo = 1
    ShanChen = 2
    ExactDifferenceMethod = 3


class Force(enum.IntEnum):
    None_ = 0
    Constant = 1
    Sinusoidal = 2
    Kolmogorov = 3
    Field = 4          # generic array read (Force.h:39-48); the array comes from mlbm_set_force_field
    ConstantShell = 5  # Force.h:296-420, 2-D lattices: synthesised on the device at mlbm_create
    EnergyRemoval = 6  # Force.h:423-561, 2-D: -amplitude x band-passed momentum of the last stored fields
    Turbulent2D = 7    # Force.h:564-616: ConstantShell + EnergyRemoval(removal_*)


class DType(enum.IntEnum):
    F64 = 0
    F32 = 1


class Overlapping(enum.IntEnum):
    Off = 0
    On = 1


LATTICE_DQ = {Lattice.D2Q5: (2, 5), Lattice.D2Q9: (2, 9), Lattice.D3Q15: (3, 15),
              Lattice.D3Q19: (3, 19), Lattice.D3Q27: (3, 27), Lattice.D2Q13: (2, 13), Lattice.D2Q17: (2, 17),
              Lattice.D2Q21: (2, 21), Lattice.D3Q33: (3, 33)}
LATTICE_INV_CS2 = {Lattice.D2Q17: 2.0 / 3.0, Lattice.D2Q21: 1.0 / (2.0 / 3.0), Lattice.D3Q33: 1.0 / 0.4156023517935171}   # else 3


class MlbmConfig(ctypes.Structure):
    """``mlbm_config`` (include/metalbm_b200.h)."""
    _fields_ = [
        ("abi_version", ctypes.c_int32),
        ("lattice", ctypes.c_int32),
        ("collision", ctypes.c_int32),
        ("equilibrium", ctypes.c_int32),
        ("forcing_scheme", ctypes.c_int32),
        ("force", ctypes.c_int32),
        ("dtype", ctypes.c_int32),
        ("overlap", ctypes.c_int32),
        ("global_length", ctypes.c_int32 * 3),
        ("rank", ctypes.c_int32),
        ("nranks", ctypes.c_int32),
        ("device", ctypes.c_int32),
        ("variant", ctypes.c_int32),
        ("tau", ctypes.c_double),
        ("force_amplitude", ctypes.c_double * 3),
        ("force_wavelength", ctypes.c_double * 3),
        ("force_k_min", ctypes.c_int32),
        ("force_k_max", ctypes.c_int32),
        ("removal_k_min", ctypes.c_int32),
        ("removal_k_max", ctypes.c_int32),
        ("removal_amplitude", ctypes.c_double * 3),
    ]


class MlbmDeviceLayout(ctypes.Structure):
    """``mlbm_device_layout`` (include/metalbm_b200.h)."""
    _fields_ = [
        ("populations", ctypes.c_void_p),
        ("component_stride", ctypes.c_size_t),
        ("plane", ctypes.c_size_t),
        ("row", ctypes.c_size_t),
        ("halo_x", ctypes.c_int32),
        ("local_length", ctypes.c_int32 * 3),
    ]


class MlbmLaunchPlan(ctypes.Structure):
    """``mlbm_launch_plan`` (include/metalbm_b200.h)."""
    _fields_ = [
        ("grid", ctypes.c_int32 * 3),
        ("block", ctypes.c_int32),
        ("shared_bytes", ctypes.c_int32),
        ("x0", ctypes.c_int32),
        ("plane_step", ctypes.c_int32),
        ("plane_count", ctypes.c_int32),
        ("planes_per_block", ctypes.c_int32),
        ("local_length", ctypes.c_int32 * 3),
        ("wrap_x", ctypes.c_int32),
        ("is_stored", ctypes.c_int32),
        ("hydro_shift", ctypes.c_int32),
        ("has_force", ctypes.c_int32),
        ("stride", ctypes.c_uint64),
        ("plane", ctypes.c_uint64),
        ("beta", ctypes.c_double),
        ("guo_factor", ctypes.c_double),
    ]


class MlbmHaloMessage(ctypes.Structure):
    """``mlbm_halo_message`` (include/metalbm_b200.h)."""
    _fields_ = [
        ("population", ctypes.c_int32),
        ("peer", ctypes.c_int32),
        ("is_send", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("offset", ctypes.c_uint64),
        ("count", ctypes.c_uint64),
    ]


def _lookup(enum_cls, value):
    if isinstance(value, enum_cls):
        return value
    if isinstance(value, str):
        name = value if value in enum_cls.__members__ else value + "_"
        return enum_cls[name]
    return enum_cls(int(value))


def make_config(lattice="D2Q9", shape=(16, 16, 1), collision="BGK", equilibrium="TruncationMa3",
                forcing_scheme="None", force="None", tau=0.7, amplitude=(0.0, 0.0, 0.0),
                wavelength=(32.0, 32.0, 32.0), dtype="F64", overlap="Off", rank=0, nranks=1,
                device=-1, variant=0, k_min=1, k_max=2, removal_amplitude=(0.0, 0.0, 0.0), removal_k_min=1,
                removal_k_max=2) -> MlbmConfig:
    cfg = MlbmConfig()
    cfg.abi_version = ABI_VERSION
    cfg.lattice = _lookup(Lattice, lattice)
    cfg.collision = _lookup(Collision, collision)
    cfg.equilibrium = _lookup(Equilibrium, equilibrium)
    cfg.forcing_scheme = _lookup(ForcingScheme, forcing_scheme)
    cfg.force = _lookup(Force, force)
    cfg.dtype = _lookup(DType, dtype)
    cfg.overlap = _lookup(Overlapping, overlap)
    dim = LATTICE_DQ[Lattice(cfg.lattice)][0]
    shape = tuple(shape) + (1,) * (3 - len(shape))
    for i in range(3):
        cfg.global_length[i] = int(shape[i]) if i < dim else 1
        cfg.force_amplitude[i] = float(amplitude[i])
        cfg.force_wavelength[i] = float(wavelength[i])
    cfg.rank, cfg.nranks, cfg.device, cfg.variant = int(rank), int(nranks), int(device), int(variant)
    cfg.tau = float(tau)
    cfg.force_k_min, cfg.force_k_max = int(k_min), int(k_max)
    cfg.removal_k_min, cfg.removal_k_max = int(removal_k_min), int(removal_k_max)
    for i in range(3):
        cfg.removal_amplitude[i] = float(removal_amplitude[i])
    return cfg


# every symbol include/metalbm_b200.h declares: name -> (restype, argtypes)
_P = ctypes.c_void_p
_SZ = ctypes.c_size_t
PROTOTYPES = {
    "mlbm_last_error": (ctypes.c_char_p, []),
    "mlbm_abi_version": (ctypes.c_int, []),
    "mlbm_create": (ctypes.c_int, [ctypes.POINTER(MlbmConfig), ctypes.POINTER(_P)]),
    "mlbm_destroy": (ctypes.c_int, [_P]),
    "mlbm_halo_plan": (ctypes.c_int, [ctypes.POINTER(MlbmConfig), ctypes.POINTER(MlbmHaloMessage), ctypes.c_int,
                                      ctypes.POINTER(ctypes.c_int)]),
    "mlbm_launch_plan_for": (ctypes.c_int, [ctypes.POINTER(MlbmConfig), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.POINTER(MlbmLaunchPlan)]),
    "mlbm_comm_unique_id": (ctypes.c_int, [_P]),
    "mlbm_comm_init": (ctypes.c_int, [_P, _P]),
    "mlbm_comm_peer_export": (ctypes.c_int, [_P, _P]),
    "mlbm_comm_peer_attach": (ctypes.c_int, [_P, _P, _P]),
    "mlbm_upload_distribution": (ctypes.c_int, [_P, _P, _SZ, _SZ, _SZ]),
    "mlbm_download_distribution": (ctypes.c_int, [_P, _P, _SZ, _SZ, _SZ]),
    "mlbm_init_equilibrium": (ctypes.c_int, [_P, _P, _P, _SZ, _SZ, _SZ]),
    "mlbm_init_synthetic": (ctypes.c_int, [_P, ctypes.c_double, ctypes.c_double]),
    "mlbm_perturb_distribution": (ctypes.c_int, [_P, ctypes.c_double, ctypes.c_uint64]),
    "mlbm_set_alpha": (ctypes.c_int, [_P, _P, _SZ, _SZ]),
    "mlbm_set_force_field": (ctypes.c_int, [_P, _P, _SZ, _SZ, _SZ]),
    "mlbm_step": (ctypes.c_int, [_P, ctypes.c_uint, ctypes.c_int]),
    "mlbm_run_async": (ctypes.c_int, [_P, ctypes.c_uint, ctypes.c_uint, ctypes.c_uint]),
    "mlbm_run_async_stored": (ctypes.c_int, [_P, ctypes.c_uint, ctypes.c_uint, ctypes.c_uint, ctypes.c_int]),
    "mlbm_sync": (ctypes.c_int, [_P]),
    "mlbm_download_fields": (ctypes.c_int, [_P, _P, _P, _P, _P, _SZ, _SZ, _SZ]),
    "mlbm_observables": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_double)]),
    "mlbm_power_spectra": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), ctypes.c_int,
                                          ctypes.POINTER(ctypes.c_int)]),
    "mlbm_alpha_statistics": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_double)]),
    "mlbm_download_halo_distribution": (ctypes.c_int, [_P, _P, _SZ]),
    "mlbm_checkpoint_write": (ctypes.c_int, [_P, ctypes.c_char_p, ctypes.c_uint]),
    "mlbm_checkpoint_read": (ctypes.c_int, [_P, ctypes.c_char_p, ctypes.POINTER(ctypes.c_uint)]),
    "mlbm_newton_statistics": (ctypes.c_int, [_P, ctypes.c_int, ctypes.POINTER(ctypes.c_ulonglong)]),
    "mlbm_reduce_sum": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_double), ctypes.c_int]),
    "mlbm_selftest_log": (ctypes.c_int, [_P, _P, _SZ]),
    "mlbm_alloc_pinned": (ctypes.c_int, [_SZ, ctypes.POINTER(_P)]),
    "mlbm_free_pinned": (ctypes.c_int, [_P]),
    "mlbm_timers": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]),
    "mlbm_device_distribution": (ctypes.c_int, [_P, ctypes.POINTER(MlbmDeviceLayout)]),
    "mlbm_launch_count": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_uint64)]),
    "mlbm_stream": (ctypes.c_int, [_P, ctypes.POINTER(_P)]),
    "mlbm_kernel_time": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint64)]),
    "mlbm_mark": (ctypes.c_int, [_P, ctypes.c_int]),
    "mlbm_elapsed": (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]),
}

_library = None


class MlbmError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"metalbm_b200 error {status}: {message}")
        self.status = status


def load_library(path: os.PathLike | None = None) -> ctypes.CDLL:
    """Load ``libmetalbm_b200.so`` (built in-tree by ``__graft_entry__.build()``) and type its symbols."""
    global _library
    if _library is not None and path is None:
        return _library
    lib_path = Path(path) if path else LIBRARY_PATH
    if not lib_path.is_file():
        raise FileNotFoundError(
            f"{lib_path} not found: build the CUDA library first (python -c 'import __graft_entry__ as g; g.build()'); "
            "there is no CPU fallback")
    lib = ctypes.CDLL(str(lib_path), mode=ctypes.RTLD_GLOBAL)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.mlbm_abi_version() != ABI_VERSION:
        raise RuntimeError("libmetalbm_b200.so ABI version mismatch")
    if path is None:
        _library = lib
    return lib


def halo_plan(cfg: MlbmConfig) -> list:
    """The halo messages of one step for ``cfg.rank`` (pure host logic, runs without a GPU)."""
    lib = load_library()
    count = ctypes.c_int()
    check(lib.mlbm_halo_plan(ctypes.byref(cfg), None, 0, ctypes.byref(count)))
    messages = (MlbmHaloMessage * max(count.value, 1))()
    check(lib.mlbm_halo_plan(ctypes.byref(cfg), messages, count.value, ctypes.byref(count)))
    return [messages[i] for i in range(count.value)]


def launch_plan(cfg: MlbmConfig, x0: int, x1: int, is_stored: int = 0, plane_step: int = 1) -> MlbmLaunchPlan:
    """Grid and scalar kernel parameters of one fused-kernel launch (pure host logic, runs without a GPU)."""
    plan = MlbmLaunchPlan()
    check(load_library().mlbm_launch_plan_for(ctypes.byref(cfg), x0, x1, is_stored, plane_step, ctypes.byref(plan)))
    return plan


def check(status: int) -> None:
    if status != 0:
        message = load_library().mlbm_last_error()
        raise MlbmError(status, message.decode() if message else "")
