"""Default integration grids and integrator selection (agnpy/utils/math.py:6-9, 55-119)."""
import functools

import numpy as np

from . import _lib

gamma_e_to_integrate = np.logspace(1, 9, 200)   # utils/math.py:6
nu_to_integrate = np.logspace(5, 30, 200)        # utils/math.py:7  [Hz]
nu_to_integrate_targets = np.logspace(3, 21, 100)  # targets/targets.py:151  [Hz]
mu_to_integrate = np.linspace(-1, 1, 100)        # utils/math.py:8
phi_to_integrate = np.linspace(0, 2 * np.pi, 50) # utils/math.py:9


# numpy.trapz was renamed numpy.trapezoid in numpy 2.0 and removed in 2.4; the reference (numpy >= 1.24,
# < 2.3) spells it np.trapz.  Resolved once here so that the package imports on either.
np_trapz = getattr(np, "trapezoid", None) or getattr(np, "trapz")
_NUMPY_TRAPZ = tuple(f for f in (getattr(np, "trapezoid", None), getattr(np, "trapz", None)) if f is not None)


def trapz_prefix(*args, **kwargs):
    """Selector for the opt-in suffix-sum evaluation of numpy.trapz over gamma in the external
    Compton kernels (AGNPY_INTEG_TRAPZ_PREFIX): same integral and masks, ~8x fewer operations."""
    raise NotImplementedError(
        "agnpy_b200.trapz_prefix only selects the device integrator; it is not a host function")


def trapz_loglog(y, x, axis=0, intervals=False):
    """Log-log trapezoidal rule of the reference, agnpy/utils/math.py:55-119: each interval is
    integrated as the power law through its end points; intervals with a zero end point or a repeated
    abscissa contribute 0.  Same signature and default axis (0) as the reference's function
    (`intervals=True` returns the per-interval integrals instead of their sum).  Quantities keep their units (y.unit * x.unit).

    Two roles: passed as `integrator=` to any evaluate_* it selects the fused device-side rule
    (AGNPY_INTEG_LOGLOG); called directly it integrates `y` along `axis` on the GPU through
    agnpy_b200_trapz_loglog_host (there is no CPU fallback: it needs a CUDA device)."""
    from . import _core
    from .units import Quantity as _Q
    y_unit = getattr(y, "unit", None)
    x_unit = getattr(x, "unit", None)
    yv = np.asarray(getattr(y, "value", y), dtype=np.float64)
    xv = np.asarray(getattr(x, "value", x), dtype=np.float64)
    if xv.ndim != 1:
        # the reference also takes x already reshaped for broadcasting against y (axes_reshaper):
        # every axis but `axis` of length 1
        if xv.ndim == yv.ndim and xv.size == xv.shape[axis]:
            xv = xv.reshape(-1)
        else:
            raise ValueError("x must be one-dimensional, or shaped like y with length 1 on every axis "
                             "but `axis`")
    if yv.shape[axis] != xv.size:
        raise ValueError(f"y has {yv.shape[axis]} samples along axis {axis}, x has {xv.size}")
    moved = np.ascontiguousarray(np.moveaxis(yv, axis, -1))
    rows = moved.reshape(-1, xv.size)
    out = _core.trapz_loglog_host(rows, xv, bool(intervals))
    out = out.reshape(moved.shape[:-1] + ((xv.size - 1,) if intervals else ()))
    if intervals:
        out = np.moveaxis(out, -1, axis)
    elif out.ndim == 0:
        out = out[()]
    if y_unit is not None or x_unit is not None:
        unit = (y_unit if y_unit is not None else 1) * (x_unit if x_unit is not None else 1)
        return out * unit
    return out


def integrator_code(integrator):
    """`integrator=` keyword -> AGNPY_INTEG_*.  Decided by identity: numpy.trapz / numpy.trapezoid, this
    package's `trapz_loglog` / `trapz_prefix`, or the strings "trapz" / "trapz_loglog" /
    "trapz_prefix".  The reference's own `agnpy.utils.math.trapz_loglog` function object is recognised
    by module + name (it cannot be imported here to compare identities).  Anything else raises --
    there is no host fallback that could run an arbitrary callable, and a user function that merely
    happens to be called `trapz` must not silently select the device rule."""
    if integrator is None or any(integrator is f for f in _NUMPY_TRAPZ):
        return _lib.INTEG_TRAPZ
    if integrator is trapz_loglog:
        return _lib.INTEG_LOGLOG
    if integrator is trapz_prefix:
        return _lib.INTEG_TRAPZ_PREFIX
    if isinstance(integrator, str):
        code = {"trapz": _lib.INTEG_TRAPZ, "trapezoid": _lib.INTEG_TRAPZ,
                "trapz_loglog": _lib.INTEG_LOGLOG, "trapz_prefix": _lib.INTEG_TRAPZ_PREFIX}.get(integrator)
        if code is not None:
            return code
    name = getattr(integrator, "__name__", "")
    module = getattr(integrator, "__module__", "") or ""
    if name == "trapz_loglog" and module == "agnpy.utils.math":
        return _lib.INTEG_LOGLOG
    if name in ("trapz", "trapezoid") and module.split(".")[0] == "numpy":
        return _lib.INTEG_TRAPZ     # e.g. numpy.lib.function_base.trapz of an older numpy
    raise ValueError(
        f"unsupported integrator {integrator!r}: use numpy.trapz / numpy.trapezoid or trapz_loglog")


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
