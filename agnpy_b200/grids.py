"""Default integration grids and integrator selection (agnpy/utils/math.py:6-9, 55-119)."""
import functools

import numpy as np

from . import _lib

gamma_e_to_integrate = np.logspace(1, 9, 200)   # utils/math.py:6
nu_to_integrate = np.logspace(5, 30, 200)        # utils/math.py:7  [Hz]
nu_to_integrate_targets = np.logspace(3, 21, 100)  # targets/targets.py:151  [Hz]
mu_to_integrate = np.linspace(-1, 1, 100)        # utils/math.py:8
phi_to_integrate = np.linspace(0, 2 * np.pi, 50) # utils/math.py:9


def trapz_prefix(*args, **kwargs):
    """Selector for the opt-in suffix-sum evaluation of numpy.trapz over gamma in the external
    Compton kernels (AGNPY_INTEG_TRAPZ_PREFIX): same integral and masks, ~8x fewer operations."""
    raise NotImplementedError(
        "agnpy_b200.trapz_prefix only selects the device integrator; it is not a host function")


def trapz_fp32(*args, **kwargs):
    """Selector for the opt-in FP32-integrand evaluation of numpy.trapz in the external Compton
    kernels (AGNPY_INTEG_TRAPZ_FP32): masks and outer sums stay FP64, rtol ~1e-6."""
    raise NotImplementedError(
        "agnpy_b200.trapz_fp32 only selects the device integrator; it is not a host function")


def trapz_loglog(*args, **kwargs):
    """Selector for the device-side log-log trapezoidal rule (utils/math.py:55-119).
    Pass it as `integrator=`; the integration itself happens inside the CUDA kernels."""
    raise NotImplementedError(
        "agnpy_b200.trapz_loglog only selects the device integrator; it is not a host function")


_TRAPZ_NAMES = {"trapz", "trapezoid"}


def integrator_code(integrator):
    """`integrator=` keyword -> AGNPY_INTEG_*.  Accepts numpy.trapz / numpy.trapezoid, the
    reference's or this package's `trapz_loglog`, or the strings "trapz" / "trapz_loglog".
    Anything else raises: there is no host fallback that could run an arbitrary callable."""
    if integrator is None:
        return _lib.INTEG_TRAPZ
    name = integrator if isinstance(integrator, str) else getattr(integrator, "__name__", "")
    if name in _TRAPZ_NAMES:
        return _lib.INTEG_TRAPZ
    if name == "trapz_loglog":
        return _lib.INTEG_LOGLOG
    if name == "trapz_prefix":
        return _lib.INTEG_TRAPZ_PREFIX
    if name == "trapz_fp32":
        return _lib.INTEG_TRAPZ_FP32
    raise ValueError(
        f"unsupported integrator {integrator!r}: use numpy.trapz or trapz_loglog")


def phi_fold_flag(phi):
    """INTEG_FOLD_PHI if the phi grid is mirror-symmetric -- phi[i] + phi[n-1-i] = 2 pi m for one
    integer m, to 1e-13 -- else 0.  True for the reference's linspace(0, 2 pi, 50): cos phi, the
    only way phi enters the Compton kernels (utils/geometry.py:5-12), then takes the same value at
    both points of a mirror pair and the EC kernels evaluate each pair once.
    (Decided once per distinct grid: a fit loop passes the same array on every call.)"""
    if phi is None:
        return 0
    p = np.ascontiguousarray(phi, dtype=np.float64)
    return _phi_fold_flag_of(p.tobytes())


@functools.lru_cache(maxsize=64)
def _phi_fold_flag_of(raw):
    p = np.frombuffer(raw, dtype=np.float64)
    if p.size < 2 or not np.all(np.isfinite(p)):
        return 0
    s = p + p[::-1]
    m = np.round(s[0] / (2 * np.pi))
    tol = 1e-13 * max(1.0, float(np.max(np.abs(p))))
    return _lib.INTEG_FOLD_PHI if np.all(np.abs(s - 2 * np.pi * m) <= tol) else 0


@functools.lru_cache(maxsize=64)
def _ascending(raw):
    a = np.frombuffer(raw, dtype=np.float64)
    return bool((a[1:] > a[:-1]).all())


def as_grid(x, name, ascending=True):
    if isinstance(x, np.ndarray) and x.dtype == np.float64 and x.ndim == 1 and x.flags.c_contiguous:
        a = x
    else:
        a = np.ascontiguousarray(np.asarray(x, dtype=np.float64).ravel())
    if a.size < 2:
        raise ValueError(f"{name}: need at least 2 grid points")
    if ascending and not _ascending(a.tobytes()):   # cached per distinct grid
        raise ValueError(f"{name}: grid must be strictly ascending")
    return a
