"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

Python face of the CPU checker: ctypes access to ``liboracle.so`` (the plain-C restatement in
``lbm_oracle.c``), the reference's *spectral* vorticity / enstrophy restated with numpy
(Transformer.h:118-295, Routine.h:129-132, Analysis.h:68-98) and the seeded synthetic initial
fields of SURVEY.md section 8(d).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference legs may
import this module.  The product package ``metalbm_b200`` never does.
"""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

from metalbm_b200.capi import LATTICE_DQ, Lattice, MlbmConfig, make_config  # config struct = the ABI's

ORACLE_DIR = Path(__file__).resolve().parent
_LIB = None


def build() -> Path:
    subprocess.run(["make", "-s", "-C", str(ORACLE_DIR)], check=True, capture_output=True)
    return ORACLE_DIR / "liboracle.so"


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        path = ORACLE_DIR / "liboracle.so"
        if not path.is_file():
            build()
        _LIB = ctypes.CDLL(str(path))
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        cp = ctypes.POINTER(MlbmConfig)
        _LIB.mlbm_oracle_step.argtypes = [cp, dp, dp, dp, dp, dp, dp, ctypes.c_int, ip, ip]
        _LIB.mlbm_oracle_step_ex.argtypes = [cp, dp, dp, dp, dp, dp, dp, ctypes.c_int, ip, ip, dp, dp]
        _LIB.mlbm_oracle_init_equilibrium.argtypes = [cp, dp, dp, dp]
        _LIB.mlbm_oracle_observables.argtypes = [cp, dp, dp, dp]
        _LIB.mlbm_oracle_lattice.argtypes = [ctypes.c_int, ip, ip, ip, dp]
    return _LIB


def _dp(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def lattice(name) -> tuple:
    """(D, Q, celerity[Q,3] int, weight[Q])."""
    lat = Lattice[name] if isinstance(name, str) else Lattice(int(name))
    d, q = ctypes.c_int(), ctypes.c_int()
    c = np.zeros((40, 3), dtype=np.int32)      # MAXQ of lbm_oracle.c is 33 (D3Q33)
    w = np.zeros(40, dtype=np.float64)
    assert lib().mlbm_oracle_lattice(int(lat), ctypes.byref(d), ctypes.byref(q), _ip(c), _dp(w)) == 0
    return d.value, q.value, c[:q.value].copy(), w[:q.value].copy()


def shape_of(cfg: MlbmConfig) -> tuple:
    dim = LATTICE_DQ[Lattice(cfg.lattice)][0]
    return tuple(int(cfg.global_length[i]) if i < dim else 1 for i in range(3))


class OracleState:
    """Global-domain state advanced by the C restatement (one ``iterate`` per ``step``)."""

    def __init__(self, cfg: MlbmConfig, populations: np.ndarray, alpha: np.ndarray | None = None):
        self.cfg = cfg
        self.shape = shape_of(cfg)
        self.dim, self.q = LATTICE_DQ[Lattice(cfg.lattice)]
        self.f = np.ascontiguousarray(populations, dtype=np.float64).reshape((self.q,) + self.shape).copy()
        self.next = np.empty_like(self.f)
        self.alpha = (np.full(self.shape, 2.0) if alpha is None
                      else np.ascontiguousarray(alpha, dtype=np.float64).reshape(self.shape).copy())
        self.density = np.zeros(self.shape)
        self.velocity = np.zeros((self.dim,) + self.shape)
        self.force = np.zeros((self.dim,) + self.shape)
        if int(cfg.force) in (5, 7):   # ConstantShell / Turbulent2D: the array setForceArray makes in the Collision ctor (Collision.h:51-54)
            self.force[...] = constant_shell_force(cfg)
        self.branch = np.zeros(self.shape, dtype=np.int32)
        self.iterations = np.zeros(self.shape, dtype=np.int32)
        # checker-side diagnostics of the last step: rounding-noise floor of the Newton iterate and max_q |fNeq_q|
        self.alpha_noise = np.zeros(self.shape)
        self.fneq_max = np.zeros(self.shape)

    def step(self, is_stored: bool = True) -> None:
        if int(self.cfg.force) in (6, 7):
            # Collision::update -> Force::update (Collision.h:97-100, Force.h:552-558, 605-609) at the top of every iterate
            # (Algorithm.h:338): the array is rebuilt from fieldList, i.e. from the fields of the LAST STORED step
            self.force[...] = time_dependent_force(self.cfg, self.density, self.velocity)
        status = lib().mlbm_oracle_step_ex(ctypes.byref(self.cfg), _dp(self.f), _dp(self.next), _dp(self.alpha),
                                           _dp(self.density), _dp(self.velocity), _dp(self.force),
                                           1 if is_stored else 0, _ip(self.branch), _ip(self.iterations),
                                           _dp(self.alpha_noise), _dp(self.fneq_max))
        if status != 0:
            raise ValueError("oracle: unsupported configuration")
        self.f, self.next = self.next, self.f

    def observables(self) -> np.ndarray:
        """[energy, enstrophy (spectral, reference definition), mach, mass] of the last stored step."""
        out = np.zeros(4)
        lib().mlbm_oracle_observables(ctypes.byref(self.cfg), _dp(self.density), _dp(self.velocity), _dp(out))
        out[1] = spectral_enstrophy(self.velocity, self.dim)
        return out


def max_wave_number(cfg: MlbmConfig) -> int:
    """gFD::maxWaveNumber() = arrayMax(globalLengthInt) / 2 (FourierDomain.h:77-79), where arrayMax_impl is
    Max(first, arrayMIN of the rest) (Helpers.h:37-39): max(Lx, min(Ly, Lz)) / 2 with the unused dimensions equal to 1."""
    nx, ny, nz = shape_of(cfg)
    return max(nx, min(ny, nz)) // 2


def power_spectra(cfg: MlbmConfig, velocity: np.ndarray, force: np.ndarray) -> np.ndarray:
    """SpectralAnalysisList::writeAnalyses (AnalysisList.h:132-170) with PowerSpectra (Analysis.h:122-177): [K, 2] energy and
    forcing spectra, K = max_wave_number (FourierDomain.h:77-79).  Per stored half-spectrum mode of the
    UNNORMALISED r2c transform: sum_d |a^_d|^2, halved where the wave number of the last dimension is 0 (:151-154, so the
    Nyquist column counts in full), added to bin floor(|k|) when that is below K (:166-167; integer wave numbers
    k = i <= N/2 ? i : i - N, kNorm = (unsigned)norm).  Only the energy spectrum is divided by the volume (normalizeAnalyses,
    AnalysisList.h:189)."""
    shape = shape_of(cfg)
    dim = LATTICE_DQ[Lattice(cfg.lattice)][0]
    lengths = shape[:dim]
    bins = max_wave_number(cfg)
    wave = [np.array([i if i <= n // 2 else i - n for i in range(n)], dtype=np.float64) for n in lengths]
    wave[-1] = np.arange(lengths[-1] // 2 + 1, dtype=np.float64)      # the halved dimension: 0 .. N/2
    grids = np.meshgrid(*wave, indexing="ij")
    k_norm = np.floor(np.sqrt(sum(g * g for g in grids))).astype(np.int64)
    coefficient = np.where(grids[-1] == 0, 0.5, 1.0)
    out = np.zeros((bins, 2))
    for column, field in enumerate((velocity, force)):
        energy = np.zeros(k_norm.shape)
        for d in range(dim):
            spectrum = np.fft.rfftn(np.asarray(field[d], dtype=np.float64).reshape(lengths))
            energy += spectrum.real ** 2 + spectrum.imag ** 2
        weighted = (coefficient * energy).ravel()
        index = k_norm.ravel()
        keep = index < bins
        out[:, column] = np.bincount(index[keep], weights=weighted[keep], minlength=bins)[:bins]
    out[:, 0] /= float(np.prod(lengths))
    return out


def constant_shell_force(cfg: MlbmConfig) -> np.ndarray:
    """Force<double, ForceType::ConstantShell> for 2-D lattices, step by step as the reference does it:
    initTempArray (Force.h:333-420): psi^ = amplitude[0] (real) on the shell kMin^2 <= |k|^2 <= kMax^2 of the r2c half
    spectrum [Nx][Ny/2+1], integer wave numbers k = i <= N/2 ? i : i - N (the mirrored writes of :353-372, :394-414 set the
    same values); MakeIncompressible<2>::executeFourier (Transformer.h:318-381): F^x = (-ky Im psi^, ky Re psi^),
    F^y = (kx Im psi^, -kx Re psi^); BackwardFFT::execute (Transformer.h:101-108): c2r, divided by the volume.
    numpy's irfft2 is that c2r (it drops the non-Hermitian part of the self-conjugate columns like FFTW) already divided
    by the volume.  Pinned against the arrays the reference itself produced (tests/golden/*constantshell*.npz)."""
    nx, ny, nz = shape_of(cfg)
    assert nz == 1 and LATTICE_DQ[Lattice(cfg.lattice)][0] == 2, "the 3-D variant corrupts the reference's heap; not restated"
    kx = np.array([i if i <= nx // 2 else i - nx for i in range(nx)], dtype=np.float64)[:, None]
    ky = np.arange(ny // 2 + 1, dtype=np.float64)[None, :]
    k2 = kx * kx + ky * ky
    psi = np.where((k2 >= cfg.force_k_min ** 2) & (k2 <= cfg.force_k_max ** 2), float(cfg.force_amplitude[0]), 0.0).astype(np.complex128)
    fx_hat = (-ky * psi.imag) + 1j * (ky * psi.real)
    fy_hat = (kx * psi.imag) + 1j * (-kx * psi.real)
    force = np.stack([np.fft.irfft2(fx_hat, s=(nx, ny)), np.fft.irfft2(fy_hat, s=(nx, ny))])
    return force.reshape(2, nx, ny, 1)


def energy_removal_force(cfg: MlbmConfig, density: np.ndarray, velocity: np.ndarray, amplitude, k_min: int, k_max: int) -> np.ndarray:
    """Force<double, ForceType::EnergyRemoval>::setForceArray (Force.h:452-550) for 2-D lattices: momentum = density *
    velocity of fieldList (:466-474), r2c (:478-480), F^_d = -amplitude[d] * momentum^_d on the shell
    kMin^2 <= |k|^2 <= kMax^2 and zero elsewhere (:485-545, integer wave numbers of the half-spectrum indices),
    c2r divided by the volume (:547-549, Transformer.h:101-108)."""
    nx, ny, nz = shape_of(cfg)
    assert nz == 1
    kx = np.array([i if i <= nx // 2 else i - nx for i in range(nx)], dtype=np.float64)[:, None]
    ky = np.arange(ny // 2 + 1, dtype=np.float64)[None, :]
    k2 = kx * kx + ky * ky
    shell = (k2 >= k_min ** 2) & (k2 <= k_max ** 2)
    force = np.zeros((2, nx, ny, 1))
    for d in range(2):
        momentum = (density * velocity[d]).reshape(nx, ny)
        spectrum = np.where(shell, -float(amplitude[d]) * np.fft.rfft2(momentum), 0.0)
        force[d, :, :, 0] = np.fft.irfft2(spectrum, s=(nx, ny))
    return force


def time_dependent_force(cfg: MlbmConfig, density: np.ndarray, velocity: np.ndarray) -> np.ndarray:
    """EnergyRemoval (Force.h:423-561) or Turbulent2D = ConstantShell injection + EnergyRemoval with the removal* constants
    (Force.h:564-616: setForceArray adds the two arrays, :583-603)."""
    if int(cfg.force) == 6:
        return energy_removal_force(cfg, density, velocity, cfg.force_amplitude, cfg.force_k_min, cfg.force_k_max)
    return constant_shell_force(cfg) + energy_removal_force(cfg, density, velocity, cfg.removal_amplitude,
                                                            cfg.removal_k_min, cfg.removal_k_max)


def init_equilibrium(cfg: MlbmConfig, density: np.ndarray, velocity: np.ndarray) -> np.ndarray:
    shape = shape_of(cfg)
    dim, q = LATTICE_DQ[Lattice(cfg.lattice)]
    rho = np.ascontiguousarray(density, dtype=np.float64).reshape(shape)
    u = np.ascontiguousarray(velocity, dtype=np.float64).reshape((dim,) + shape)
    f = np.empty((q,) + shape)
    assert lib().mlbm_oracle_init_equilibrium(ctypes.byref(cfg), _dp(rho), _dp(u), _dp(f)) == 0
    return f


# ---------------------------------------------------------------------------------------------
# The reference's vorticity and enstrophy (spectral, integer wavenumbers, double normalisation).
#   Curl<...,2,2>  Transformer.h:118-187     Curl<...,3,3>  Transformer.h:189-295
#   BackwardFFT::execute divides by V (Transformer.h:101-108); Curl::normalize divides by V again
#   (Transformer.h:179-185, 284-294); TotalEnstrophy = sum 0.5*w^2 / V (Analysis.h:85-93, :30).
# ---------------------------------------------------------------------------------------------
def _wavenumbers(n: int, half: bool) -> np.ndarray:
    count = n // 2 + 1 if half else n
    i = np.arange(count)
    return np.where(i <= n // 2, i, i - n).astype(np.float64)


def spectral_vorticity(velocity: np.ndarray, dim: int) -> np.ndarray:
    """velocity [D, nx, ny, nz] -> vorticity [2D-3, nx, ny, nz] exactly as the reference stores it."""
    shape = velocity.shape[1:]
    volume = float(np.prod(shape))
    if dim == 2:
        nx, ny = shape[0], shape[1]
        ux = np.fft.rfft2(velocity[0, :, :, 0])
        uy = np.fft.rfft2(velocity[1, :, :, 0])
        kx = _wavenumbers(nx, False)[:, None]
        ky = _wavenumbers(ny, True)[None, :]
        # only the real part is formed, the imaginary part is forced to zero (Transformer.h:151-165)
        w_hat = (-kx * uy.imag + ky * ux.imag).astype(np.complex128)
        w = np.fft.irfft2(w_hat, s=(nx, ny)) * (nx * ny)  # FFTW c2r is unnormalised
        return (w / volume / volume).reshape((1,) + shape)
    nx, ny, nz = shape
    u = [np.fft.rfftn(velocity[d]) for d in range(3)]
    kx = _wavenumbers(nx, False)[:, None, None]
    ky = _wavenumbers(ny, False)[None, :, None]
    kz = _wavenumbers(nz, True)[None, None, :]
    wx = 1j * (ky * u[2] - kz * u[1])
    wy = 1j * (kz * u[0] - kx * u[2])
    wz = 1j * (kx * u[1] - ky * u[0])
    out = np.stack([np.fft.irfftn(w, s=shape, axes=(0, 1, 2)) * volume for w in (wx, wy, wz)])
    return out / volume / volume


def spectral_enstrophy(velocity: np.ndarray, dim: int) -> float:
    w = spectral_vorticity(velocity, dim)
    return float((0.5 * w * w).sum() / np.prod(velocity.shape[1:]))


# ---------------------------------------------------------------------------------------------
# Seeded synthetic inputs ("Init B"): f = feq(rho, u) * (1 + eps * N(0,1))
# ---------------------------------------------------------------------------------------------
def synthetic_populations(cfg: MlbmConfig, eps: float = 1e-2, seed: int = 20261017,
                          amplitude: float = 0.05, ripple: float = 0.05) -> np.ndarray:
    shape = shape_of(cfg)
    dim, q = LATTICE_DQ[Lattice(cfg.lattice)]
    nx, ny, nz = shape
    x = (2 * np.pi * np.arange(nx) / nx)[:, None, None]
    y = (2 * np.pi * np.arange(ny) / ny)[None, :, None]
    z = (2 * np.pi * np.arange(nz) / nz)[None, None, :]
    rho = 1.0 + ripple * np.sin(x) * np.cos(y) * np.cos(z) * np.ones(shape)
    u = np.zeros((dim,) + shape)
    if dim == 2:
        u[0] = amplitude * np.sin(y) * np.ones(shape)
        u[1] = amplitude * np.cos(x) * np.ones(shape)
    else:
        u[0] = amplitude * np.sin(x) * np.cos(y) * np.cos(z)
        u[1] = -amplitude * np.cos(x) * np.sin(y) * np.cos(z)
        u[2] = 0.5 * amplitude * np.cos(x) * np.cos(y) * np.sin(z)
    f = init_equilibrium(cfg, rho, u)
    rng = np.random.default_rng(seed)
    return f * (1.0 + eps * rng.standard_normal(f.shape))
