"""Flat LambdaCDM luminosity distance with astropy's treatment of photons and massive neutrinos
(Komatsu et al. 2011 fitting formula) — stub, test infrastructure only.  `Planck18` is astropy's default
cosmology; pinned by the reference's emission_regions/tests/test_emission_regions.py:44-50."""
import numpy as np
from scipy.integrate import quad

from .. import units as u


class FlatLambdaCDM:
    def __init__(self, H0, Om0, Tcmb0=0.0, Neff=3.04, m_nu=(0.0, 0.0, 0.0), Ob0=None, name=None):
        self.name = name
        self.H0 = float(H0)             # km / s / Mpc
        self.Om0 = float(Om0)
        self.Tcmb0 = float(Tcmb0)       # K
        self.Neff = float(Neff)
        m_nu = np.atleast_1d(np.asarray(m_nu, dtype=float))     # eV
        h = self.H0 / 100.0
        # critical density today, g / cm3
        G = 6.6743e-8
        H0_s = self.H0 * 1e5 / 3.0856775814913673e24
        rho_c = 3.0 * H0_s ** 2 / (8.0 * np.pi * G)
        # photons: a T^4 / c^2
        a_B = 4.0 * 5.670374419e-5 / 2.99792458e10
        self.Ogamma0 = a_B * self.Tcmb0 ** 4 / 2.99792458e10 ** 2 / rho_c
        self.Tnu0 = 0.7137658555036082 * self.Tcmb0
        self._massive = m_nu[m_nu > 0]
        self._nmassless = int(np.sum(m_nu == 0))
        nnu = len(m_nu)
        self._neff_per_nu = self.Neff / nnu if nnu else 0.0
        kB_eV = 8.617333262145179e-5
        self._nu_y = self._massive / (kB_eV * self.Tnu0) if self.Tnu0 > 0 else self._massive * 0
        self.Onu0 = self.Ogamma0 * self.nu_relative_density(0.0)
        self.Ode0 = 1.0 - self.Om0 - self.Ogamma0 - self.Onu0

    def nu_relative_density(self, z):
        prefac = 0.22710731766          # 7/8 (4/11)^(4/3)
        if len(self._massive) == 0:
            return prefac * self.Neff
        p, invp, k = 1.83, 0.54644808743, 0.3173
        curr = self._nu_y / (1.0 + z)
        rel_mass = np.sum((1.0 + (k * curr) ** p) ** invp) + self._nmassless
        return prefac * self._neff_per_nu * rel_mass

    def inv_efunc(self, z):
        zp1 = 1.0 + z
        Or = self.Ogamma0 * (1.0 + self.nu_relative_density(z))
        return (zp1 ** 3 * (Or * zp1 + self.Om0) + self.Ode0) ** -0.5

    def luminosity_distance(self, z):
        def one(zz):
            integral = quad(self.inv_efunc, 0.0, zz, epsabs=0, epsrel=1e-12)[0]
            return (1.0 + zz) * 299792.458 / self.H0 * integral
        out = np.vectorize(one)(np.asarray(z, dtype=float))
        return u.Quantity(out, u.Mpc)


Planck18 = FlatLambdaCDM(H0=67.66, Om0=0.30966, Tcmb0=2.7255, Neff=3.046, m_nu=(0.0, 0.0, 0.06),
                         Ob0=0.04897, name="Planck18")
Planck15 = FlatLambdaCDM(H0=67.74, Om0=0.3075, Tcmb0=2.7255, Neff=3.046, m_nu=(0.0, 0.0, 0.06),
                         Ob0=0.0486, name="Planck15")
WMAP9 = FlatLambdaCDM(H0=69.32, Om0=0.2865, Tcmb0=2.725, Neff=3.04, m_nu=(0.0, 0.0, 0.0),
                      Ob0=0.04628, name="WMAP9")
default_cosmology = Planck18
