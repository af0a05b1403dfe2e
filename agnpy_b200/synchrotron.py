"""Drop-in for agnpy/synchrotron/synchrotron.py: `Synchrotron` with the reference's static
`evaluate_sed_flux` / `evaluate_tau_ssa` signatures (synchrotron.py:101-144), running on the
fused FP64 CUDA kernel K1."""
import numpy as np

from . import _core, _lib, spectra
from . import constants as const
from .grids import as_grid, gamma_e_to_integrate, integrator_code, np_trapz
from .units import strip, wrap

B_CR = 4.414e13   # critical magnetic field in G (synchrotron.py:13), delta approximation only

__all__ = ["Synchrotron", "ProtonSynchrotron"]


def _blob_models(z, d_L, delta_D, B, R_b, ne_par):
    models = _lib.make_models(1)
    _core.model_row(models, 0, z=float(z), d_L=strip(d_L, "cm", "d_L"), delta_D=float(delta_D),
                    B=strip(B, "G", "B"), R_b=strip(R_b, "cm", "R_b"), ne=ne_par)
    return models


def _ne_setup(n_e, args, gamma, ssa):
    """kind, parameters, the gamma grid to upload and (for interpolated n_e) the tables"""
    kind, par = spectra.describe(n_e, args)
    gamma = as_grid(gamma, "gamma")
    ne_table = ssa_table = None
    if kind == _lib.NE_TABLE:
        g = spectra.snap_boundaries(gamma, par[1], par[2])
        ne_table, ssa_table = spectra.tables(n_e, args, g, ssa)
    return kind, par, gamma, ne_table, ssa_table


class Synchrotron:
    """synchrotron.py:71-277.  `blob` is any object exposing z, d_L, delta_D, B, R_b, n_e
    (with `.parameters`) and gamma_e, e.g. an agnpy Blob."""

    def __init__(self, blob, ssa=False, integrator=np_trapz):
        self.blob = blob
        self.ssa = ssa
        self.integrator = integrator

    @staticmethod
    def evaluate_tau_ssa(nu, z, d_L, delta_D, B, R_b, n_e, *args, integrator=np_trapz,
                         gamma=gamma_e_to_integrate):
        """SSA optical depth, synchrotron.py:101-129; returns a plain ndarray."""
        nu_hz = strip(nu, "Hz", "nu")
        kind, par, gamma, ne_t, ssa_t = _ne_setup(n_e, args, gamma, True)
        models = _blob_models(z, d_L, delta_D, B, R_b, par)
        _, tau = _core.synch_sed_host(models, kind, np.ravel(nu_hz), gamma, ne_t, ssa_t, True,
                                      integrator_code(integrator), want_tau=True)
        return tau[0].reshape(np.shape(nu_hz))

    @staticmethod
    def evaluate_sed_flux(nu, z, d_L, delta_D, B, R_b, n_e, *args, ssa=False,
                          integrator=np_trapz, gamma=gamma_e_to_integrate):
        """nu F_nu [erg cm-2 s-1] of synchrotron radiation, synchrotron.py:131-211."""
        nu_hz = strip(nu, "Hz", "nu")
        kind, par, gamma, ne_t, ssa_t = _ne_setup(n_e, args, gamma, ssa)
        models = _blob_models(z, d_L, delta_D, B, R_b, par)
        sed = _core.synch_sed_host(models, kind, np.ravel(nu_hz), gamma, ne_t, ssa_t, ssa,
                                   integrator_code(integrator))
        return wrap(sed[0].reshape(np.shape(nu_hz)), "erg cm-2 s-1")

    def electron_energy_loss_rate(self, gamma):
        """synchrotron.py:93-99: 4/3 sigma_T c U_B gamma^2 [erg s-1] (closed form, host)"""
        from . import constants as const
        B = strip(self.blob.B, "G", "B")
        U_B = B**2 / (8 * np.pi)
        return wrap((4 / 3) * const.sigma_T * const.c * U_B * np.asarray(gamma, dtype=np.float64) ** 2,
                    "erg s-1")

    def sed_flux(self, nu):
        """synchrotron.py:229-244"""
        b = self.blob
        return self.evaluate_sed_flux(nu, b.z, b.d_L, b.delta_D, b.B, b.R_b, b.n_e,
                                      *b.n_e.parameters, ssa=self.ssa,
                                      integrator=self.integrator, gamma=b.gamma_e)

    @staticmethod
    def evaluate_sed_flux_delta_approx(nu, z, d_L, delta_D, B, R_b, n_e, *args):
        """delta-function approximation of the synchrotron SED, Dermer & Menon (2009) eq. 7.70
        (synchrotron.py:213-227): a closed form per frequency -- no integral, hence host arithmetic:
        each frequency maps to one Lorentz factor gamma_s = sqrt(eps' B_cr / B), and the SED is
        delta_D^4 c sigma_T U_B gamma_s^3 V_b n_e(gamma_s) / (6 pi d_L^2)."""
        nu_hz = strip(nu, "Hz", "nu")
        d_L, B, R_b = strip(d_L, "cm", "d_L"), strip(B, "G", "B"), strip(R_b, "cm", "R_b")
        eps = (1 + z) * (const.h * nu_hz / const.mec2) / delta_D
        gamma_s = np.sqrt(eps / (B / B_CR))
        n = n_e.evaluate(gamma_s, *args)
        n = strip(n, "cm-3", "n_e") if hasattr(n, "to_value") else np.asarray(n, dtype=np.float64)
        V_b = 4 / 3 * np.pi * R_b**3
        U_B = B**2 / (8 * np.pi)
        pref = delta_D**4 * const.c * const.sigma_T * U_B / (6 * np.pi * d_L**2)
        return wrap(pref * gamma_s**3 * (V_b * n), "erg cm-2 s-1")

    def sed_flux_delta_approx(self, nu):
        """synchrotron.py:246-258"""
        b = self.blob
        return self.evaluate_sed_flux_delta_approx(nu, b.z, b.d_L, b.delta_D, b.B, b.R_b, b.n_e,
                                                   *b.n_e.parameters)

    def sed_luminosity(self, nu):
        """synchrotron.py:259-265"""
        d_L = strip(self.blob.d_L, "cm", "d_L")
        return wrap(4 * np.pi * d_L**2 * strip(self.sed_flux(nu), "erg cm-2 s-1"), "erg s-1")

    def sed_peak_flux(self, nu):
        return self.sed_flux(nu).max()

    def sed_peak_nu(self, nu):
        return nu[self.sed_flux(nu).argmax()]


class ProtonSynchrotron:
    """synchrotron/proton_synchrotron.py:14-187: synchrotron radiation of the protons of a blob
    (`blob.n_p`, `blob.gamma_p`); the reference has no self-absorption for protons."""

    def __init__(self, blob, integrator=np_trapz):
        self.blob = blob
        self.integrator = integrator

    @staticmethod
    def evaluate_sed_flux(nu, z, d_L, delta_D, B, R_b, n_p, *args, integrator=np_trapz,
                          gamma=gamma_e_to_integrate):
        """proton_synchrotron.py:69-151"""
        nu_hz = strip(nu, "Hz", "nu")
        kind, par, gamma, ne_t, _ = _ne_setup(n_p, args, gamma, False)
        models = _blob_models(z, d_L, delta_D, B, R_b, par)
        sed = _core.proton_synch_sed_host(models, kind, np.ravel(nu_hz), gamma, ne_t,
                                          integrator_code(integrator))
        return wrap(sed[0].reshape(np.shape(nu_hz)), "erg cm-2 s-1")

    def sed_flux(self, nu):
        """proton_synchrotron.py:153-167"""
        b = self.blob
        return self.evaluate_sed_flux(nu, b.z, b.d_L, b.delta_D, b.B, b.R_b, b.n_p,
                                      *b.n_p.parameters, integrator=self.integrator, gamma=b.gamma_p)

    def sed_luminosity(self, nu):
        d_L = strip(self.blob.d_L, "cm", "d_L")
        return wrap(4 * np.pi * d_L**2 * strip(self.sed_flux(nu), "erg cm-2 s-1"), "erg s-1")

    def sed_peak_flux(self, nu):
        return self.sed_flux(nu).max()

    def sed_peak_nu(self, nu):
        return nu[self.sed_flux(nu).argmax()]
