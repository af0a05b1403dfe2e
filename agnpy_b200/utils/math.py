"""Names of agnpy/utils/math.py at the reference's import path (`from agnpy.utils.math import ...`).

The default integration grids and `trapz_loglog` are the ones the drop-in classes use
(agnpy_b200/grids.py; `trapz_loglog` integrates on the device).  `axes_reshaper` and `log` are the
reference's small host helpers for user code that builds its own integrands: the kernels never
materialise those broadcast grids."""
import numpy as np

from ..grids import (gamma_e_to_integrate, mu_to_integrate, nu_to_integrate, phi_to_integrate,  # noqa: F401
                     trapz_loglog)

numpy_type = np.float64
ftiny = np.finfo(numpy_type).tiny   # utils/math.py:17
fmax = np.finfo(numpy_type).max     # utils/math.py:19
min_rel_distance = 1.0e-4           # utils/math.py:12


def axes_reshaper(*args):
    """k one-dimensional arrays -> k views broadcastable against each other: array i keeps its length
    on axis i and has length 1 elsewhere (utils/math.py:22-46)"""
    n = len(args)
    out = []
    for i, arg in enumerate(args):
        shape = [1] * n
        shape[i] = np.size(arg)
        out.append(np.reshape(arg, shape))
    return out


def log(x):
    """logarithm that never returns -inf or nan: the argument is clipped to [tiny, max] (utils/math.py:49-52)"""
    return np.log(np.clip(np.asarray(getattr(x, "value", x), dtype=numpy_type), ftiny, fmax))


def power_law_integral(k_e, p, gamma_min, gamma_max):
    """integral of k_e gamma^-p between gamma_min and gamma_max (utils/math.py:121-129)"""
    if np.isclose(p, 1.0):
        return k_e * np.log(gamma_max / gamma_min)
    return k_e * (gamma_max ** (1 - p) - gamma_min ** (1 - p)) / (1 - p)
