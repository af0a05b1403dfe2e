"""Host-side luminosity distance d_L(z) for the callers that start from a redshift.

The reference obtains d_L from astropy on every model evaluation — `Distance(z=z).to("cm")` in the fit
wrappers (agnpy/fit/sherpa_wrapper.py:63, 106, 180, 259; gammapy_wrapper.py:51), in `Blob.__init__`
(emission_regions/blob.py:81) and in the targets' `sed_flux` (targets/targets.py:424, 614) — with astropy's
default cosmology, Planck18: flat LambdaCDM, H0 = 67.66 km/s/Mpc, Om0 = 0.30966, Tcmb0 = 2.7255 K,
Neff = 3.046, one massive neutrino of 0.06 eV (treated with the Komatsu et al. 2011 fitting formula, as
astropy does).  The kernels take d_L, never z -> d_L; this helper is scalar host arithmetic outside the
hot path and caches per redshift (a fit evaluates thousands of vectors at one or a few z).

Known answers of the reference (emission_regions/tests/test_emission_regions.py:44-50):
    d_L(2.1) = 5.21497473e28 cm,  d_L(0.1) = 1.4682341e27 cm   (reproduced to all printed digits).
"""
import functools

import numpy as np

from . import constants as const

__all__ = ["FlatLambdaCDM", "Planck18", "luminosity_distance"]

_MPC_CM = 3.0856775814913673e24
_SIGMA_SB = 5.670374419e-5          # astropy CODATA 2018 literal, erg cm-2 s-1 K-4
_KB_EV = 8.617333262145179e-5       # k_B in eV / K
_GL_X, _GL_W = np.polynomial.legendre.leggauss(48)


class FlatLambdaCDM:
    """flat LambdaCDM with photons and (massive) neutrinos; luminosity distance only"""

    def __init__(self, H0, Om0, Tcmb0=0.0, Neff=3.04, m_nu=(0.0, 0.0, 0.0), name=None):
        self.name, self.H0, self.Om0, self.Tcmb0, self.Neff = name, float(H0), float(Om0), float(Tcmb0), float(Neff)
        m_nu = np.atleast_1d(np.asarray(m_nu, dtype=np.float64))
        H0_cgs = self.H0 * 1e5 / _MPC_CM
        rho_crit = 3.0 * H0_cgs**2 / (8.0 * np.pi * const.G)
        self.Ogamma0 = 4.0 * _SIGMA_SB / const.c**3 * self.Tcmb0**4 / rho_crit
        Tnu0 = 0.7137658555036082 * self.Tcmb0          # (4/11)^(1/3) Tcmb0
        massive = m_nu[m_nu > 0]
        self._n_massless = int(np.count_nonzero(m_nu == 0))
        self._neff_per_nu = self.Neff / m_nu.size if m_nu.size else 0.0
        self._nu_y = massive / (_KB_EV * Tnu0) if Tnu0 > 0 else np.zeros(0)
        self.Onu0 = self.Ogamma0 * self._nu_relative_density(np.zeros(1))[0]
        self.Ode0 = 1.0 - self.Om0 - self.Ogamma0 - self.Onu0
        self._cached = functools.lru_cache(maxsize=4096)(self._d_L_cm)

    def _nu_relative_density(self, z):
        """neutrino / photon energy density at redshift z (Komatsu et al. 2011, eq. 26)"""
        prefac = 0.22710731766          # 7/8 (4/11)^(4/3)
        if self._nu_y.size == 0:
            return np.full_like(z, prefac * self.Neff)
        p, invp, k = 1.83, 0.54644808743, 0.3173
        curr = self._nu_y[None, :] / (1.0 + z[:, None])
        rel_mass = np.sum((1.0 + (k * curr) ** p) ** invp, axis=1) + self._n_massless
        return prefac * self._neff_per_nu * rel_mass

    def inv_efunc(self, z):
        z = np.asarray(z, dtype=np.float64)
        zp1 = 1.0 + z
        Or = self.Ogamma0 * (1.0 + self._nu_relative_density(np.ravel(z)).reshape(z.shape))
        return (zp1**3 * (Or * zp1 + self.Om0) + self.Ode0) ** -0.5

    def _d_L_cm(self, z):
        if z < 0:
            raise ValueError("redshift must be non-negative")
        # 48-point Gauss-Legendre on panels of at most 0.25 in z: the integrand is smooth, the
        # quadrature error is below 1e-15
        n_panels = max(1, int(np.ceil(z / 0.25)))
        edges = np.linspace(0.0, z, n_panels + 1)
        half = 0.5 * (edges[1:] - edges[:-1])
        mid = 0.5 * (edges[1:] + edges[:-1])
        zz = mid[:, None] + half[:, None] * _GL_X[None, :]
        integral = float(np.sum(half[:, None] * _GL_W[None, :] * self.inv_efunc(zz)))
        d_H = const.c / (self.H0 * 1e5) * _MPC_CM       # Hubble distance, cm
        return (1.0 + z) * d_H * integral

    def luminosity_distance(self, z):
        """d_L in cm (float for a scalar z, ndarray otherwise)"""
        if np.ndim(z) == 0:
            return self._cached(float(z))
        z = np.asarray(z, dtype=np.float64)
        uniq, inv = np.unique(z, return_inverse=True)
        return np.array([self._cached(float(v)) for v in uniq])[inv].reshape(z.shape)


# astropy.cosmology.Planck18 (astropy's default cosmology since 5.0)
Planck18 = FlatLambdaCDM(H0=67.66, Om0=0.30966, Tcmb0=2.7255, Neff=3.046, m_nu=(0.0, 0.0, 0.06),
                         name="Planck18")


def luminosity_distance(z, cosmology=None):
    """`Distance(z=z).to_value("cm")` of the reference: d_L in cm, cached per redshift"""
    return (Planck18 if cosmology is None else cosmology).luminosity_distance(z)
