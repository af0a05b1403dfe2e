"""Names of agnpy/utils/conversion.py at the reference's import path, in CGS: the quantities carry the
package's units (astropy's when it is installed, the CGS stand-in otherwise)."""
from .. import constants as const
from ..units import Quantity, strip, wrap

mec2 = wrap(const.mec2, "erg")                # conversion.py:6
mpc2 = wrap(const.m_p * const.c**2, "erg")    # conversion.py:7
lambda_c_e = wrap(const.lambda_c_e, "cm")     # conversion.py:8
lambda_c_p = wrap(const.h / (const.m_p * const.c), "cm")  # conversion.py:9
Gauss_cgs_unit = "cm(-1/2) g(1/2) s-1"        # conversion.py:11


def nu_to_epsilon_prime(nu, z=0, delta_D=1, m=None):
    """frequency -> dimensionless energy h nu / (m c^2) in a frame at redshift z moving with Doppler factor
    delta_D (conversion.py:30-35); m: particle mass in g (default: the electron)"""
    mass = const.m_e if m is None else strip(m, "g", "m")
    return (1 + z) * (const.h * strip(nu, "Hz", "nu") / (mass * const.c**2)) / delta_D


def B_to_cgs(B):
    """magnetic field in Gaussian-CGS base units (conversion.py:38-40)"""
    return wrap(strip(B, "G", "B"), Gauss_cgs_unit)


def to_R_g_units(r, M):
    """distance in units of the gravitational radius G M / c^2 (conversion.py:43-46)"""
    return strip(r, "cm", "r") / (const.G * strip(M, "g", "M") / const.c**2)


__all__ = ["mec2", "mpc2", "lambda_c_e", "lambda_c_p", "Gauss_cgs_unit", "nu_to_epsilon_prime", "B_to_cgs",
           "to_R_g_units", "Quantity"]
