"""Parameter containers with the attribute names of agnpy/targets/targets.py and
agnpy/emission_regions/blob.py, for environments without agnpy/astropy.  Only what the hot
path reads is kept (energy densities, black-body SEDs, printing and plotting stay with the
reference)."""
import numpy as np

from . import _core, _lib, constants as const, cosmology as _cosmology
from .grids import nu_to_integrate_targets
from .units import strip, u, wrap

__all__ = ["CMB", "PointSourceBehindJet", "SSDisk", "SphericalShellBLR", "RingDustTorus",
           "Blob", "lines_dictionary"]

# broad emission lines of the stratified BLR model (Finke 2016, table 5; the reference's
# targets.py:22-150 holds the same published numbers)
_LINES = (  # name, rest wavelength [Angstrom], R_line / R_Hbeta, L_line / L_Hbeta
    ('Lyepsilon', 937.8, 2.7, 0.24),
    ('Lydelta', 949.74, 2.8, 0.24),
    ('CIII', 977.02, 0.83, 0.6),
    ('NIII', 990.69, 0.85, 0.6),
    ('Lybeta', 1025.72, 1.2, 1.1),
    ('OVI', 1033.83, 1.2, 1.1),
    ('ArI', 1066.66, 4.5, 0.094),
    ('Lyalpha', 1215.67, 0.27, 12),
    ('OI', 1304.35, 4.0, 0.23),
    ('SiII', 1306.82, 4.0, 0.23),
    ('SiIV', 1396.76, 0.83, 1.0),
    ('OIV]', 1402.06, 0.83, 1.0),
    ('CIV', 1549.06, 0.83, 2.9),
    ('NIV', 1718.55, 3.8, 0.3),
    ('AlII', 1721.89, 3.8, 0.3),
    ('CIII]', 1908.73, 0.46, 1.8),
    ('[NeIV]', 2423.83, 5.8, 0.051),
    ('MgII', 2798.75, 0.45, 1.7),
    ('Hdelta', 4102.89, 3.4, 0.12),
    ('Hgamma', 4341.68, 3.2, 0.3),
    ('HeII', 4687.02, 0.63, 0.016),
    ('Hbeta', 4862.68, 1.0, 1.0),
    ('[ClIII]', 5539.43, 4.8, 0.039),
    ('HeI', 5877.29, 0.39, 0.092),
    ('Halpha', 6564.61, 1.3, 3.6),
)


class _LineDictionary(dict):
    """`lines_dictionary[name]` -> {"lambda": Quantity [Angstrom], "R_Hbeta_ratio", "L_Hbeta_ratio"} as in
    the reference; the wavelength is wrapped on access so that it follows the active units module"""

    def __getitem__(self, name):
        lam, r_ratio, l_ratio = dict.__getitem__(self, name)
        return {"lambda": wrap(lam, "Angstrom"), "R_Hbeta_ratio": r_ratio, "L_Hbeta_ratio": l_ratio}

    def wavelength(self, name):
        """rest wavelength in Angstrom as a plain number"""
        return dict.__getitem__(self, name)[0]


lines_dictionary = _LineDictionary({name: (lam, r, l) for name, lam, r, l in _LINES})


def _d_L_of(z, d_L=None, cosmology=None):
    """d_L as a Quantity in cm: the given one, else the reference's `Distance(z=z, cosmology=...)`
    (emission_regions/blob.py:81, targets.py:424, 614) through agnpy_b200.cosmology (Planck18) or any
    object with astropy's `luminosity_distance(z)`"""
    if d_L is not None:
        return d_L
    if cosmology is None or isinstance(cosmology, _cosmology.FlatLambdaCDM):
        return _cosmology.luminosity_distance(float(z), cosmology) * u.cm
    return cosmology.luminosity_distance(z)


def _u_call(variant, tgt, r, blob, n_mu=50):
    """u(r[, blob]) of one target through agnpy_b200_target_u (targets.py:202-218, 429-466,
    516-543, 620-638); r keeps its shape, the result is a Quantity in erg cm-3"""
    r_cm = strip(r, "cm", "r")
    models = _lib.make_models(1)
    _core.model_row(models, 0, tgt=tgt)
    Gamma = None if blob is None else float(blob.Gamma)
    out = _core.target_u_host(variant, models, np.ravel(r_cm), Gamma, n_mu)
    return wrap(out[0].reshape(np.shape(r_cm)), "erg cm-3")


class CMB:
    """targets.py:154-183"""

    def __init__(self, z):
        T = 2.72548
        self.T, self.z = wrap(T, "K"), z
        self.u_0 = wrap((7.5657 * 1e-15 * np.power(T, 4)) * np.power(1 + z, 4), "erg cm-3")
        self.epsilon_0 = (2.7 * const.k_B * T / const.mec2) * (1 + z)

    def __str__(self):
        return f"* CMB radiation field at z = {self.z}: u_0 = {strip(self.u_0, 'erg cm-3'):.2e} erg cm-3, " \
               f"epsilon_0 = {self.epsilon_0:.2e}"

    def u(self, blob=None):
        """targets.py:171-183 (two flops: evaluated on the host)"""
        if blob:
            return self.u_0 * (np.power(blob.Gamma, 2) * (1 + np.power(blob.Beta, 2) / 3))
        return self.u_0


class PointSourceBehindJet:
    """targets.py:186-204"""

    def __init__(self, L_0, epsilon_0):
        self.L_0, self.epsilon_0 = L_0, epsilon_0

    def __str__(self):
        return f"* point source behind the jet: L_0 = {strip(self.L_0, 'erg s-1'):.2e} erg s-1, " \
               f"epsilon_0 = {self.epsilon_0:.2e}"

    def u(self, r, blob=None):
        """targets.py:202-218"""
        return _u_call(_lib.EC_PS_BEHIND_JET, [self.epsilon_0, strip(self.L_0, "erg s-1", "L_0")], r, blob)


class SSDisk:
    """targets.py:221-276"""

    def __init__(self, M_BH, L_disk, eta, R_in, R_out, R_g_units=False):
        self.name = "Shakura Sunyaev Accretion Disk"
        self.M_BH, self.L_disk, self.eta = M_BH, L_disk, eta
        M, L = strip(M_BH, "g", "M_BH"), strip(L_disk, "erg s-1", "L_disk")
        # targets.py:240-248: Eddington luminosity 1.26e46 M_8 erg/s, accretion rate L / (eta c^2)
        self.M_8 = M / (1e8 * const.M_sun)
        self.L_Edd = wrap(1.26 * 1e46 * self.M_8, "erg s-1")
        self.l_Edd = L / (1.26 * 1e46 * self.M_8)
        self.m_dot = wrap(L / (eta * const.c**2), "g s-1")
        R_g = const.G * M / const.c**2
        self.R_g = wrap(R_g, "cm")
        numbers = (int, float, np.integer, np.floating)
        if R_g_units:      # targets.py:251-260: plain numbers in units of R_g
            if not isinstance(R_in, numbers) or not isinstance(R_out, numbers):
                raise TypeError("R_in / R_out passed with units, int / float expected")
            self.R_in_tilde, self.R_out_tilde = R_in, R_out
            self.R_in, self.R_out = wrap(R_in * R_g, "cm"), wrap(R_out * R_g, "cm")
        else:              # targets.py:261-271: quantities
            if not hasattr(R_in, "to_value") or not hasattr(R_out, "to_value"):
                raise TypeError("R_in / R_out passed without units")
            self.R_in, self.R_out = R_in, R_out
            self.R_in_tilde = strip(R_in, "cm", "R_in") / R_g
            self.R_out_tilde = strip(R_out, "cm", "R_out") / R_g
        self.R_tilde = np.linspace(self.R_in_tilde, self.R_out_tilde)

    def __str__(self):
        return (f"* Shakura-Sunyaev accretion disk: M_BH = {strip(self.M_BH, 'g'):.2e} g, "
                f"L_disk = {strip(self.L_disk, 'erg s-1'):.2e} erg s-1, eta = {self.eta:.2e}, "
                f"m_dot = {strip(self.m_dot, 'g s-1'):.2e} g s-1, R_in = {strip(self.R_in, 'cm'):.2e} cm, "
                f"R_out = {strip(self.R_out, 'cm'):.2e} cm")

    # ---- small host helpers of the reference's class (targets.py:289-348), used by its callers and tests
    @staticmethod
    def evaluate_mu_from_r_tilde(R_in_tilde, R_out_tilde, r_tilde, size=100):
        """cosines under which the disk is seen from the height r_tilde (Finke 2016, eqs. 72-73)"""
        mu_min = 1 / np.sqrt(1 + np.power(R_out_tilde / r_tilde, 2))
        mu_max = 1 / np.sqrt(1 + np.power(R_in_tilde / r_tilde, 2))
        return np.linspace(mu_min, mu_max, size)

    @staticmethod
    def evaluate_phi_disk_mu(mu, R_in_tilde, r_tilde):
        """1 - sqrt(R_in / R) at the radius seen under the cosine mu (Dermer 2009, eq. 63)"""
        return 1 - np.sqrt(R_in_tilde / (r_tilde * np.sqrt(np.power(mu, -2) - 1)))

    @staticmethod
    def evaluate_epsilon(L_disk, M_BH, eta, R_tilde):
        """dimensionless energy of the photons emitted at R_tilde (Dermer 2009, eq. 65)"""
        M_8 = strip(M_BH, "g", "M_BH") / (1e8 * const.M_sun)
        l_Edd = strip(L_disk, "erg s-1", "L_disk") / (1.26 * 1e46 * M_8)
        return 2.7 * 1e-4 * np.power(l_Edd / (M_8 * eta), 1 / 4) * np.power(R_tilde, -3 / 4)

    @staticmethod
    def evaluate_epsilon_mu(L_disk, M_BH, eta, mu, r_tilde):
        return SSDisk.evaluate_epsilon(L_disk, M_BH, eta, r_tilde * np.sqrt(np.power(mu, -2) - 1))

    @staticmethod
    def evaluate_T(R, M_BH, m_dot, R_in):
        """disk temperature at the radius R (Dermer 2009, eq. 64)"""
        R_cm, Rin = strip(R, "cm", "R"), strip(R_in, "cm", "R_in")
        phi = 1 - np.sqrt(Rin / R_cm)
        val = (3 * const.G * strip(M_BH, "g", "M_BH") * strip(m_dot, "g s-1", "m_dot") * phi) / (
            8 * np.pi * np.power(R_cm, 3) * const.sigma_sb)
        return wrap(np.power(val, 1 / 4), "K")

    def epsilon_mu(self, mu, r_tilde):
        return self.evaluate_epsilon_mu(self.L_disk, self.M_BH, self.eta, mu, r_tilde)

    def epsilon(self, R_tilde):
        return self.evaluate_epsilon(self.L_disk, self.M_BH, self.eta, R_tilde)

    def phi_disk_mu(self, mu, r_tilde):
        return self.evaluate_phi_disk_mu(mu, self.R_in_tilde, r_tilde)

    def phi_disk(self, R_tilde):
        return 1 - np.sqrt(self.R_in_tilde / R_tilde)

    def T(self, R):
        return self.evaluate_T(R, self.M_BH, self.m_dot, self.R_in)

    @staticmethod
    def _bb(variant, nu, z, M_BH, m_dot, R_in, R_out, d_L, mu_s, L_disk=None):
        nu_hz = strip(nu, "Hz", "nu")
        models = _lib.make_models(1)
        tgt = [strip(M_BH, "g", "M_BH"), strip(m_dot, "g s-1", "m_dot"), strip(R_in, "cm", "R_in"),
               strip(R_out, "cm", "R_out")]
        if L_disk is not None:
            tgt.append(strip(L_disk, "erg s-1", "L_disk"))
        _core.model_row(models, 0, z=float(z), d_L=strip(d_L, "cm", "d_L"), mu_s=float(mu_s), tgt=tgt)
        sed = _core.bb_sed_host(variant, models, np.ravel(nu_hz), nu_to_integrate_targets, 100)
        return wrap(sed[0].reshape(np.shape(nu_hz)), "erg cm-2 s-1")

    @staticmethod
    def evaluate_multi_T_bb_sed(nu, z, M_BH, m_dot, R_in, R_out, d_L, mu_s=1):
        """multi-temperature black-body SED of the disk, targets.py:350-391"""
        return SSDisk._bb(_lib.BB_DISK, nu, z, M_BH, m_dot, R_in, R_out, d_L, mu_s)

    @staticmethod
    def evaluate_multi_T_bb_norm_sed(nu, z, L_disk, M_BH, m_dot, R_in, R_out, d_L, mu_s=1):
        """same, normalised to an integral luminosity L_disk, targets.py:393-415"""
        return SSDisk._bb(_lib.BB_DISK_NORM, nu, z, M_BH, m_dot, R_in, R_out, d_L, mu_s, L_disk)

    def sed_flux(self, nu, z, mu_s=1, d_L=None):
        """targets.py:417-427; d_L = Distance(z) (Planck18) unless given"""
        d_L = _d_L_of(z, d_L)
        return self.evaluate_multi_T_bb_norm_sed(nu, z, self.L_disk, self.M_BH, self.m_dot, self.R_in,
                                                 self.R_out, d_L, mu_s)

    def u(self, r, blob=None):
        """targets.py:429-466"""
        tgt = [strip(self.M_BH, "g", "M_BH"), strip(self.L_disk, "erg s-1", "L_disk"), float(self.eta),
               strip(self.R_in, "cm", "R_in"), strip(self.R_out, "cm", "R_out")]
        return _u_call(_lib.EC_SS_DISK, tgt, r, blob, n_mu=100)


class SphericalShellBLR:
    """targets.py:469-497"""

    def __init__(self, L_disk, xi_line, line, R_line):
        if line not in lines_dictionary:
            raise NameError(f"{line} not available in the line dictionary")
        self.L_disk, self.xi_line, self.line, self.R_line = L_disk, xi_line, line, R_line
        self.lambda_line = wrap(lines_dictionary.wavelength(line), "Angstrom")
        lam_cm = lines_dictionary.wavelength(line) * 1e-8
        self.epsilon_line = const.h * const.c / lam_cm / const.mec2

    def __str__(self):
        return (f"* spherical-shell broad line region: L_disk = {strip(self.L_disk, 'erg s-1'):.2e} erg s-1, "
                f"xi_line = {self.xi_line:.2e}, line = {self.line} "
                f"({strip(self.lambda_line, 'Angstrom'):.2f} Angstrom), R_line = {strip(self.R_line, 'cm'):.2e} cm")

    def u(self, r, blob=None):
        """targets.py:516-543"""
        tgt = [strip(self.L_disk, "erg s-1", "L_disk"), float(self.xi_line), self.epsilon_line,
               strip(self.R_line, "cm", "R_line")]
        return _u_call(_lib.EC_BLR, tgt, r, blob, n_mu=50)


class RingDustTorus:
    """targets.py:546-582"""

    def __init__(self, L_disk, xi_dt, T_dt, R_dt=None):
        self.L_disk, self.xi_dt, self.T_dt = L_disk, xi_dt, T_dt
        T = strip(T_dt, "K", "T_dt")
        self.Theta = const.k_B * T / const.mec2
        self.epsilon_dt = 2.7 * self.Theta
        if R_dt is None:
            L = strip(L_disk, "erg s-1", "L_disk")
            self.R_dt = wrap(3.5 * 1e18 * np.sqrt(L / 1e45) * np.power(T / 1e3, -2.6), "cm")
        else:
            self.R_dt = R_dt

    def __str__(self):
        return (f"* ring dust torus: L_disk = {strip(self.L_disk, 'erg s-1'):.2e} erg s-1, xi_dt = {self.xi_dt:.2e}, "
                f"T_dt = {strip(self.T_dt, 'K'):.2e} K, R_dt = {strip(self.R_dt, 'cm'):.2e} cm")

    @staticmethod
    def _bb(variant, nu, z, T_dt, R_dt, d_L, L_dt=None):
        nu_hz = strip(nu, "Hz", "nu")
        models = _lib.make_models(1)
        tgt = [strip(T_dt, "K", "T_dt"), strip(R_dt, "cm", "R_dt")]
        if L_dt is not None:
            tgt.append(strip(L_dt, "erg s-1", "L_dt"))
        _core.model_row(models, 0, z=float(z), d_L=strip(d_L, "cm", "d_L"), tgt=tgt)
        sed = _core.bb_sed_host(variant, models, np.ravel(nu_hz), None, 2)
        return wrap(sed[0].reshape(np.shape(nu_hz)), "erg cm-2 s-1")

    @staticmethod
    def evaluate_bb_sed(nu, z, T_dt, R_dt, d_L):
        """black-body SED of the torus, targets.py:593-600"""
        return RingDustTorus._bb(_lib.BB_DT, nu, z, T_dt, R_dt, d_L)

    @staticmethod
    def evaluate_bb_norm_sed(nu, z, L_dt, T_dt, R_dt, d_L):
        """same, normalised to the torus luminosity, targets.py:602-610"""
        return RingDustTorus._bb(_lib.BB_DT_NORM, nu, z, T_dt, R_dt, d_L, L_dt)

    def sed_flux(self, nu, z, d_L=None):
        """targets.py:612-618; d_L = Distance(z) (Planck18) unless given"""
        d_L = _d_L_of(z, d_L)
        L_dt = wrap(self.xi_dt * strip(self.L_disk, "erg s-1", "L_disk"), "erg s-1")
        return self.evaluate_bb_norm_sed(nu, z, L_dt, self.T_dt, self.R_dt, d_L)

    def u(self, r, blob=None):
        """targets.py:620-638"""
        tgt = [strip(self.L_disk, "erg s-1", "L_disk"), float(self.xi_dt), self.epsilon_dt,
               strip(self.R_dt, "cm", "R_dt")]
        return _u_call(_lib.EC_DT, tgt, r, blob)


class Blob:
    """emission_regions/blob.py:22-151: same constructor as the reference (d_L from z through the
    cosmology, Planck18 by default), plus an optional explicit `d_L` as the last keyword."""

    def __init__(self, R_b=1e16 * u.cm, z=0.069, delta_D=10, Gamma=10, B=1 * u.G, n_e=None, n_p=None,
                 xi=1.0, gamma_e_size=200, gamma_p_size=200, cosmology=None, d_L=None):
        if not isinstance(delta_D, (int, float, np.integer, np.floating)) or delta_D <= 0:
            raise ValueError("delta_D must be a positive number")
        from .spectra import PowerLaw
        self.R_b, self.z, self.delta_D, self.Gamma, self.B = R_b, z, delta_D, Gamma, B
        self.d_L = _d_L_of(z, d_L, cosmology)
        self.n_e = PowerLaw() if n_e is None else n_e
        self.n_p, self.xi = n_p, xi
        self.gamma_e_size, self.gamma_p_size = gamma_e_size, gamma_p_size

    @property
    def V_b(self):
        return wrap(4 / 3 * np.pi * np.power(strip(self.R_b, "cm", "R_b"), 3), "cm3")

    @property
    def t_var(self):
        """blob.py:99-104, in days like the reference"""
        t = ((1 + self.z) * strip(self.R_b, "cm", "R_b")) / (const.c * self.delta_D)
        return wrap(t / 86400.0, "d")

    @property
    def gamma_p(self):
        """blob.py:153-163"""
        return np.logspace(np.log10(self.n_p.gamma_min), np.log10(self.n_p.gamma_max),
                           self.gamma_p_size)

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
