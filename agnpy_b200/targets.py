"""Parameter containers with the attribute names of agnpy/targets/targets.py and
agnpy/emission_regions/blob.py, for environments without agnpy/astropy.  Only what the hot
path reads is kept (energy densities, black-body SEDs, printing and plotting stay with the
reference)."""
import numpy as np

from . import constants as const
from .units import strip, u

__all__ = ["CMB", "PointSourceBehindJet", "SSDisk", "SphericalShellBLR", "RingDustTorus",
           "Blob", "lines_dictionary"]

# rest wavelengths [Angstrom] of a few broad lines (Finke 2016, table 5)
lines_dictionary = {"Lyalpha": 1215.67, "Lybeta": 1025.72, "CIV": 1549.06, "MgII": 2797.92,
                    "Hbeta": 4861.33, "Halpha": 6562.80}


class CMB:
    """targets.py:154-169"""

    def __init__(self, z):
        T = 2.72548
        self.u_0 = (7.5657 * 1e-15 * np.power(T, 4)) * np.power(1 + z, 4) * u.Unit("erg cm-3")
        self.epsilon_0 = (2.7 * const.k_B * T / const.mec2) * (1 + z)


class PointSourceBehindJet:
    """targets.py:186-204"""

    def __init__(self, L_0, epsilon_0):
        self.L_0, self.epsilon_0 = L_0, epsilon_0


class SSDisk:
    """targets.py:221-276"""

    def __init__(self, M_BH, L_disk, eta, R_in, R_out, R_g_units=False):
        self.M_BH, self.L_disk, self.eta = M_BH, L_disk, eta
        R_g = const.G * strip(M_BH, "g", "M_BH") / const.c**2
        self.R_g = R_g * u.cm
        if R_g_units:
            self.R_in_tilde, self.R_out_tilde = float(R_in), float(R_out)
            self.R_in, self.R_out = R_in * R_g * u.cm, R_out * R_g * u.cm
        else:
            self.R_in, self.R_out = R_in, R_out
            self.R_in_tilde = strip(R_in, "cm", "R_in") / R_g
            self.R_out_tilde = strip(R_out, "cm", "R_out") / R_g


class SphericalShellBLR:
    """targets.py:469-497"""

    def __init__(self, L_disk, xi_line, line, R_line):
        if line not in lines_dictionary:
            raise NameError(f"{line} not available in the line dictionary")
        self.L_disk, self.xi_line, self.line, self.R_line = L_disk, xi_line, line, R_line
        lam_cm = lines_dictionary[line] * 1e-8
        self.epsilon_line = const.h * const.c / lam_cm / const.mec2


class RingDustTorus:
    """targets.py:546-582"""

    def __init__(self, L_disk, xi_dt, T_dt, R_dt=None):
        self.L_disk, self.xi_dt, self.T_dt = L_disk, xi_dt, T_dt
        T = strip(T_dt, "K", "T_dt")
        self.Theta = const.k_B * T / const.mec2
        self.epsilon_dt = 2.7 * self.Theta
        if R_dt is None:
            L = strip(L_disk, "erg s-1", "L_disk")
            self.R_dt = 3.5 * 1e18 * np.sqrt(L / 1e45) * np.power(T / 1e3, -2.6) * u.cm
        else:
            self.R_dt = R_dt


class Blob:
    """emission_regions/blob.py:61-151; d_L must be given (the reference derives it from z
    with astropy's cosmology)."""

    def __init__(self, R_b, z, d_L, delta_D, Gamma, B, n_e, gamma_e_size=200):
        if delta_D <= 0:
            raise ValueError("delta_D must be a positive number")
        self.R_b, self.z, self.d_L, self.delta_D, self.Gamma, self.B = R_b, z, d_L, delta_D, Gamma, B
        self.n_e = n_e
        self.gamma_e_size = gamma_e_size

    @property
    def Beta(self):
        return np.sqrt(1 - 1 / np.power(self.Gamma, 2))

    @property
    def mu_s(self):
        return (1 - 1 / (self.Gamma * self.delta_D)) / self.Beta

    @property
    def gamma_e(self):
        return np.logspace(np.log10(self.n_e.gamma_min), np.log10(self.n_e.gamma_max),
                           self.gamma_e_size)

    @property
    def gamma_e_external_frame(self):
        return np.logspace(1, 9, self.gamma_e_size)
