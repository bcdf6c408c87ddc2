"""Host-side mirror of the reference's step-driver interface for this path, over the C-ABI.

Names and call order follow the reference so that the parity tests read like ``Routine::compute``
(Routine.h:90-154):

    distribution = Distribution(domain)           # Distribution.h:15-43 (host local-padded SoA)
    fieldList    = FieldList(domain)              # FieldList.h:22-63
    algorithm    = Algorithm(config, fieldList, distribution, communication)   # Algorithm.h:317-324
    algorithm.unpack()                            # Algorithm.h:141-147
    algorithm.isStored = ...; algorithm.iterate(iteration)                     # Algorithm.h:326-358
    algorithm.pack()                              # Algorithm.h:132-139

Everything numerical happens in ``libmetalbm_b200.so`` (hand-written CUDA); this module only owns
numpy host arrays in the reference's layouts and forwards calls.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import capi
from .capi import LATTICE_DQ, DType, Lattice, MlbmConfig, check, load_library


class Domain:
    """lSD / gSD index arithmetic (Domain.h:42-171): local lengths, FFTW-style padding of the last used
    dimension (``ProjectPadRealAndLeave1``, MathVector.h:330-344) and ``numberElements`` (FFTWInitializer.h:22-26)."""

    def __init__(self, config: MlbmConfig):
        self.dim, self.q = LATTICE_DQ[Lattice(config.lattice)]
        self.global_length = tuple(int(config.global_length[i]) if i < self.dim else 1 for i in range(3))
        self.rank, self.nranks = int(config.rank), int(config.nranks)
        if self.global_length[0] % self.nranks:
            raise ValueError("numProcs must divide globalLengthX (Domain.h:22-24)")
        self.local_length = (self.global_length[0] // self.nranks,) + self.global_length[1:]
        padded = list(self.local_length)
        last = self.dim - 1
        padded[last] = 2 * (self.local_length[last] // 2 + 1)
        self.padded_length = tuple(padded)                       # lSD::pLength (Domain.h:53-57)
        self.number_elements = int(np.prod(self.padded_length))  # FFTWInit::numberElements == lSD::pVolume
        self.dtype = np.float64 if config.dtype == DType.F64 else np.float32

    @property
    def offset_x(self) -> int:
        """gSD::sOffset(rank)[d::X] (Domain.h:155-162)."""
        return self.rank * self.local_length[0]

    def allocate(self, components: int) -> np.ndarray:
        return np.zeros((components,) + self.padded_length, dtype=self.dtype)

    def interior(self, array: np.ndarray) -> np.ndarray:
        lx, ly, lz = self.local_length
        return array[..., :lx, :ly, :lz]


class Distribution:
    """Host side of ``Distribution<T, Architecture::GPU>`` (Distribution.h:15-43): the local padded SoA array the
    reference checkpoints and initialises; the ping-pong halo pair lives on the device inside the context."""

    def __init__(self, domain: Domain, allocate: bool = True):
        self.domain = domain
        self.array = domain.allocate(domain.q) if allocate else None

    def set_interior(self, populations: np.ndarray) -> None:
        """populations: [Q, lx, ly, lz] of this rank's slab."""
        self.domain.interior(self.array)[...] = populations

    def get_interior(self) -> np.ndarray:
        return np.ascontiguousarray(self.domain.interior(self.array))


class FieldList:
    """density, velocity, alpha, force (FieldList.h:22-63) in the reference's local padded layout."""

    def __init__(self, domain: Domain, allocate: bool = True):
        self.domain = domain
        if not allocate:   # slabs that fill the GPU: the host never sees the fields
            self.density = self.velocity = self.alpha = self.force = None
            return
        self.density = domain.allocate(1)
        self.velocity = domain.allocate(domain.dim)
        self.alpha = domain.allocate(1)
        self.alpha[...] = 2.0                      # initAlpha (Initialize.h:82-88)
        self.force = domain.allocate(domain.dim)


class Communication:
    """Replaces ``MPIInitializer`` + ``Communication`` (MPIInitializer.h:27-58, Communication.h:104-222) for one
    rank: ring neighbours and the NCCL id hand-shake.  ``broadcast`` ships 128 bytes from rank 0 to all ranks;
    by default it uses ``torch.distributed`` (any backend, gloo works)."""

    def __init__(self, rank: int = 0, nranks: int = 1, broadcast=None, all_gather=None):
        self.rank, self.nranks = rank, nranks
        self._all_gather = all_gather
        self.rank_left = (rank + nranks - 1) % nranks    # MPIInitializer.h:56
        self.rank_right = (rank + 1) % nranks            # MPIInitializer.h:57
        self._broadcast = broadcast

    def broadcast_bytes(self, payload: bytes | None, size: int) -> bytes:
        if self.nranks == 1:
            return payload
        if self._broadcast is not None:
            return self._broadcast(payload, size)
        import torch
        import torch.distributed as dist
        buffer = torch.zeros(size, dtype=torch.uint8)
        if self.rank == 0:
            buffer = torch.frombuffer(bytearray(payload), dtype=torch.uint8).clone()
        if dist.get_backend() == "nccl":
            buffer = buffer.cuda()
        dist.broadcast(buffer, src=0)
        return bytes(buffer.cpu().numpy().tobytes())

    def all_gather_bytes(self, payload: bytes) -> list:
        """Every rank's ``payload`` (equal sizes) on every rank: ships the CUDA IPC handles of the direct peer halos."""
        if self.nranks == 1:
            return [payload]
        if self._all_gather is not None:
            return self._all_gather(payload)
        import torch
        import torch.distributed as dist
        mine = torch.frombuffer(bytearray(payload), dtype=torch.uint8).clone()
        if dist.get_backend() == "nccl":
            mine = mine.cuda()
        everyone = [torch.zeros_like(mine) for _ in range(self.nranks)]
        dist.all_gather(everyone, mine)
        return [bytes(t.cpu().numpy().tobytes()) for t in everyone]


def slab_of(global_array: np.ndarray, rank: int, nranks: int) -> np.ndarray:
    """x-slab of rank ``rank`` of a [..., nx, ny, nz] global array (Domain.h:22-24)."""
    nx = global_array.shape[-3]
    lx = nx // nranks
    return global_array[..., rank * lx:(rank + 1) * lx, :, :]


class Algorithm:
    """``Algorithm<T, Pull, GPU, SoA, OneD, MPI, Overlapping>`` (Algorithm.h:300-452) over the C-ABI."""

    def __init__(self, config: MlbmConfig, field_list: FieldList | None = None,
                 distribution: Distribution | None = None, communication: Communication | None = None,
                 host_distribution: bool = True, peer_halos: bool | None = None, host_fields: bool = True):
        self._lib = load_library()
        self.config = config
        self.domain = Domain(config)
        self.fieldList = field_list if field_list is not None else FieldList(self.domain, allocate=host_fields)
        self.distribution = (distribution if distribution is not None
                             else Distribution(self.domain, allocate=host_distribution))
        self.communication = communication or Communication(int(config.rank), int(config.nranks))
        self.isStored = False
        self.peer_halos = False
        self._ctx = ctypes.c_void_p()
        check(self._lib.mlbm_create(ctypes.byref(config), ctypes.byref(self._ctx)))
        if config.nranks > 1:
            unique = ctypes.create_string_buffer(128)
            if config.rank == 0:
                check(self._lib.mlbm_comm_unique_id(unique))
            payload = self.communication.broadcast_bytes(unique.raw if config.rank == 0 else None, 128)
            check(self._lib.mlbm_comm_init(self._ctx, ctypes.create_string_buffer(payload, 128)))
            # direct peer halos (boundary kernel stores into the neighbours' halo planes over NVLink): default on with
            # Overlapping::On; MLBM_PEER_HALOS=0 or peer_halos=False keeps the NCCL send/recv exchange
            if peer_halos is None:
                single_speed = Lattice(config.lattice) in (Lattice.D2Q5, Lattice.D2Q9, Lattice.D3Q15, Lattice.D3Q19, Lattice.D3Q27)
                peer_halos = int(config.overlap) == 1 and single_speed and os.environ.get("MLBM_PEER_HALOS", "1") != "0"
            self.peer_halos = False
            if peer_halos:
                self._attach_peers()

    def _attach_peers(self) -> None:
        mine = ctypes.create_string_buffer(capi.PEER_HANDLE_BYTES)
        check(self._lib.mlbm_comm_peer_export(self._ctx, mine))
        blobs = self.communication.all_gather_bytes(mine.raw)
        left = ctypes.create_string_buffer(blobs[self.communication.rank_left], capi.PEER_HANDLE_BYTES)
        right = ctypes.create_string_buffer(blobs[self.communication.rank_right], capi.PEER_HANDLE_BYTES)
        check(self._lib.mlbm_comm_peer_attach(self._ctx, left, right))
        self.peer_halos = True

    # -- life cycle -------------------------------------------------------------------------------
    def close(self) -> None:
        if self._ctx:
            self._lib.mlbm_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- Algorithm::unpack / pack -----------------------------------------------------------------
    def _layout(self):
        d = self.domain
        return d.number_elements, d.padded_length[1], d.padded_length[2]

    def unpack(self) -> None:
        stride, py, pz = self._layout()
        check(self._lib.mlbm_upload_distribution(self._ctx, self.distribution.array.ctypes.data, stride, py, pz))

    def pack(self) -> None:
        stride, py, pz = self._layout()
        check(self._lib.mlbm_download_distribution(self._ctx, self.distribution.array.ctypes.data, stride, py, pz))

    # -- DistributionWriter / DistributionReader (Writer.h:400-445, Reader.h:119-157) ----------------
    def write_checkpoint(self, path, iteration: int = 0) -> None:
        """This rank's hyperslab of the dimQ "distribution<iQ>" data sets (padded global box, doubles) into `path`."""
        check(self._lib.mlbm_checkpoint_write(self._ctx, os.fsencode(str(path)), int(iteration)))

    def read_checkpoint(self, path) -> int:
        """The inverse: loads this rank's hyperslab, returns the iteration the file was written at."""
        iteration = ctypes.c_uint(0)
        check(self._lib.mlbm_checkpoint_read(self._ctx, os.fsencode(str(path)), ctypes.byref(iteration)))
        return int(iteration.value)

    def init_equilibrium(self) -> None:
        """initDistribution (Initialize.h:106-117) from fieldList.density / velocity, on the device."""
        stride, py, pz = self._layout()
        check(self._lib.mlbm_init_equilibrium(self._ctx, self.fieldList.density.ctypes.data,
                                              self.fieldList.velocity.ctypes.data, stride, py, pz))

    def init_synthetic(self, density_amplitude: float = 0.05, velocity_amplitude: float = 0.05) -> None:
        """The benchmark's synthetic initial field (SURVEY 8d "Init B") evaluated on the device, no host arrays."""
        check(self._lib.mlbm_init_synthetic(self._ctx, float(density_amplitude), float(velocity_amplitude)))

    def perturb(self, eps: float, seed: int = 20261017) -> None:
        """f *= 1 + eps * noise on the device (synthetic non-equilibrium fields, decomposition independent)."""
        check(self._lib.mlbm_perturb_distribution(self._ctx, float(eps), int(seed)))

    def set_alpha(self) -> None:
        _, py, pz = self._layout()
        check(self._lib.mlbm_set_alpha(self._ctx, self.fieldList.alpha.ctypes.data, py, pz))

    def set_force(self) -> None:
        """Force::update / setForceArray (Force.h:51-54, 323-331): hand ``fieldList.force`` to the device as the array the
        generic force read uses (contexts created with force="Field")."""
        stride, py, pz = self._layout()
        check(self._lib.mlbm_set_force_field(self._ctx, self.fieldList.force.ctypes.data, stride, py, pz))

    # -- Algorithm::iterate -----------------------------------------------------------------------
    def iterate(self, iteration: int) -> None:
        check(self._lib.mlbm_step(self._ctx, iteration, 1 if self.isStored else 0))
        if self.isStored and self.fieldList.density is not None:
            self.fetch_fields()

    def run(self, first_iteration: int, count: int, store_every: int = 0, sync: bool = True, stored_mode: int = 1) -> None:
        check(self._lib.mlbm_run_async_stored(self._ctx, first_iteration, count, store_every, stored_mode))
        if sync:
            self.synchronize()

    def synchronize(self) -> None:
        check(self._lib.mlbm_sync(self._ctx))

    def fetch_fields(self) -> None:
        """Bring the stored fields into ``fieldList`` (the reference writes them to pinned host memory from the
        kernel, Field.h:62-79; here they stay on the device until asked for)."""
        stride, py, pz = self._layout()
        f = self.fieldList
        check(self._lib.mlbm_download_fields(self._ctx, f.density.ctypes.data, f.velocity.ctypes.data,
                                             f.alpha.ctypes.data, f.force.ctypes.data, stride, py, pz))

    def observables(self) -> np.ndarray:
        """[total energy, total enstrophy, max Mach, total mass] of the last stored step, reduced over ranks."""
        out = (ctypes.c_double * 4)()
        check(self._lib.mlbm_observables(self._ctx, out))
        return np.array(list(out))

    def power_spectra(self) -> np.ndarray:
        """[K, 2]: energy spectrum of the stored velocity and forcing spectrum of the force array of the last stored step
        (SpectralAnalysisList::writeAnalyses, AnalysisList.h:132-170), reduced over ranks; K = gFD::maxWaveNumber()."""
        count = ctypes.c_int()
        check(self._lib.mlbm_power_spectra(self._ctx, None, None, 0, ctypes.byref(count)))
        energy, forcing = (ctypes.c_double * max(count.value, 1))(), (ctypes.c_double * max(count.value, 1))()
        check(self._lib.mlbm_power_spectra(self._ctx, energy, forcing, count.value, ctypes.byref(count)))
        return np.stack([np.array(energy[:count.value]), np.array(forcing[:count.value])], axis=1)

    def alpha_statistics(self) -> tuple:
        """(fraction of nodes whose alpha left the shortcut value 2, min alpha, max alpha) of the last step, over all ranks."""
        out = (ctypes.c_double * 3)()
        check(self._lib.mlbm_alpha_statistics(self._ctx, out))
        return float(out[0]), float(out[1]), float(out[2])

    def newton_statistics(self, start: bool = False) -> tuple:
        """start=True: begin counting; else (nodes that took the Newton solve, evaluations of F and F') on this rank since then."""
        out = (ctypes.c_ulonglong * 2)()
        check(self._lib.mlbm_newton_statistics(self._ctx, 1 if start else 0, out))
        return int(out[0]), int(out[1])

    def getCommunicationTime(self) -> float:
        c, _ = ctypes.c_double(), ctypes.c_double()
        check(self._lib.mlbm_timers(self._ctx, ctypes.byref(c), None))
        return c.value

    def getComputationTime(self) -> float:
        c = ctypes.c_double()
        check(self._lib.mlbm_timers(self._ctx, None, ctypes.byref(c)))
        return c.value

    # -- bench support ----------------------------------------------------------------------------
    def launch_count(self) -> int:
        n = ctypes.c_uint64()
        check(self._lib.mlbm_launch_count(self._ctx, ctypes.byref(n)))
        return n.value

    def kernel_time(self) -> tuple:
        ms, n = ctypes.c_double(), ctypes.c_uint64()
        check(self._lib.mlbm_kernel_time(self._ctx, ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    def mark(self, slot: int) -> None:
        check(self._lib.mlbm_mark(self._ctx, slot))

    def elapsed_ms(self, start: int, stop: int) -> float:
        ms = ctypes.c_double()
        check(self._lib.mlbm_elapsed(self._ctx, start, stop, ctypes.byref(ms)))
        return ms.value

    def stream(self) -> int:
        s = ctypes.c_void_p()
        check(self._lib.mlbm_stream(self._ctx, ctypes.byref(s)))
        return s.value or 0

    def device_layout(self) -> capi.MlbmDeviceLayout:
        layout = capi.MlbmDeviceLayout()
        check(self._lib.mlbm_device_distribution(self._ctx, ctypes.byref(layout)))
        return layout
