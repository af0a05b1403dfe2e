"""Drop-in for agnpy/compton: `SynchrotronSelfCompton` (synchrotron_self_compton.py:73-158)
and `ExternalCompton` (external_compton.py:63-634) with the reference's static signatures,
running on the fused FP64 CUDA kernels K2 / K3."""
import numpy as np

from . import _core, _lib, spectra
from .grids import (as_grid, gamma_e_to_integrate, integrator_code, mu_to_integrate, np_trapz,
                    nu_to_integrate, phi_fold_flag, phi_to_integrate)
from .synchrotron import _blob_models, _ne_setup
from .units import strip, wrap

__all__ = ["SynchrotronSelfCompton", "ExternalCompton"]


class SynchrotronSelfCompton:
    """synchrotron_self_compton.py:19-193"""

    def __init__(self, blob, ssa=False, integrator=np_trapz):
        self.blob = blob
        self.ssa = ssa
        self.integrator = integrator

    @staticmethod
    def evaluate_sed_flux(nu, z, d_L, delta_D, B, R_b, n_e, *args, ssa=False,
                          integrator=np_trapz, gamma=gamma_e_to_integrate):
        """nu F_nu [erg cm-2 s-1] of SSC, synchrotron_self_compton.py:73-158; the seed
        synchrotron spectrum is evaluated at utils/math.py:7 `nu_to_integrate`."""
        nu_hz = strip(nu, "Hz", "nu")
        kind, par, gamma, ne_t, ssa_t = _ne_setup(n_e, args, gamma, ssa)
        models = _blob_models(z, d_L, delta_D, B, R_b, par)
        sed = _core.ssc_sed_host(models, kind, np.ravel(nu_hz), nu_to_integrate, gamma, ne_t,
                                 ssa_t, ssa, integrator_code(integrator))
        return wrap(sed[0].reshape(np.shape(nu_hz)), "erg cm-2 s-1")

    def electron_energy_loss_rate(self, gamma):
        """Klein-Nishina corrected SSC loss rate [erg s-1] at the Lorentz factors `gamma`,
        synchrotron_self_compton.py:43-71 (seed spectrum on utils/math.py:6-7 default grids)"""
        b = self.blob
        kind, par = spectra.describe(b.n_e, b.n_e.parameters)
        grid = as_grid(gamma_e_to_integrate, "gamma")
        ne_t = None
        if kind == _lib.NE_TABLE:
            ne_t, _ = spectra.tables(b.n_e, b.n_e.parameters,
                                     spectra.snap_boundaries(grid, par[1], par[2]), False)
        models = _blob_models(b.z, b.d_L, b.delta_D, b.B, b.R_b, par)
        g = np.asarray(gamma, dtype=np.float64)
        out = _core.ssc_loss_rate_host(models, kind, np.ravel(g), nu_to_integrate, grid, ne_t)
        return wrap(out[0].reshape(g.shape), "erg s-1")

    def sed_flux(self, nu):
        """synchrotron_self_compton.py:160-175"""
        b = self.blob
        return self.evaluate_sed_flux(nu, b.z, b.d_L, b.delta_D, b.B, b.R_b, b.n_e,
                                      *b.n_e.parameters, ssa=self.ssa,
                                      integrator=self.integrator, gamma=b.gamma_e)

    def sed_luminosity(self, nu):
        d_L = strip(self.blob.d_L, "cm", "d_L")
        return wrap(4 * np.pi * d_L**2 * strip(self.sed_flux(nu), "erg cm-2 s-1"), "erg s-1")

    def sed_peak_flux(self, nu):
        return self.sed_flux(nu).max()

    def sed_peak_nu(self, nu):
        return nu[self.sed_flux(nu).argmax()]


def _ec_call(variant, nu, z, d_L, delta_D, mu_s, R_b, tgt, n_e, args, integrator, gamma, mu,
             n_mu, phi):
    nu_hz = strip(nu, "Hz", "nu")
    kind, par = spectra.describe(n_e, args)
    gamma = as_grid(gamma, "gamma")
    ne_table = None
    if kind == _lib.NE_TABLE:  # n_e at snap(gamma/delta_D), external_compton.py:125-126
        g = spectra.snap_boundaries(gamma / float(delta_D), par[1], par[2])
        ne_table, _ = spectra.tables(n_e, args, g, False)
    models = _lib.make_models(1)
    _core.model_row(models, 0, z=float(z), d_L=strip(d_L, "cm", "d_L"), delta_D=float(delta_D),
                    mu_s=float(mu_s), R_b=strip(R_b, "cm", "R_b"), ne=par, tgt=tgt)
    sed = _core.ec_sed_host(variant, models, kind, np.ravel(nu_hz), gamma, mu, n_mu, phi,
                            ne_table, integrator_code(integrator) | phi_fold_flag(phi))
    return wrap(sed[0].reshape(np.shape(nu_hz)), "erg cm-2 s-1")


def _target_kind(target):
    name = type(target).__name__
    return name if name in ("CMB", "PointSourceBehindJet", "SSDisk", "SphericalShellBLR",
                            "RingDustTorus") else getattr(target, "kind", name)


class ExternalCompton:
    """external_compton.py:25-640.  `target` is an agnpy target (or an object of
    agnpy_b200.targets) recognised by class name."""

    def __init__(self, blob, target, r=None, integrator=np_trapz):
        self.blob = blob
        self.target = target
        self.r = r
        self.integrator = integrator
        self.set_mu()
        self.set_phi()

    def set_mu(self, mu_size=100):
        """external_compton.py:48-57"""
        self.mu_size = mu_size
        if _target_kind(self.target) == "SSDisk":
            t = self.target
            r_tilde = strip(self.r, "cm", "r") / strip(t.R_g, "cm", "R_g")
            mu_min = 1 / np.sqrt(1 + np.power(t.R_out_tilde / r_tilde, 2))
            mu_max = 1 / np.sqrt(1 + np.power(t.R_in_tilde / r_tilde, 2))
            self.mu = np.linspace(mu_min, mu_max, 100)
        else:
            self.mu = np.linspace(-1, 1, self.mu_size)

    def set_phi(self, phi_size=50):
        """external_compton.py:59-61"""
        self.phi_size = phi_size
        self.phi = np.linspace(0, 2 * np.pi, self.phi_size)

    # ---- static evaluators, same signatures as the reference ---------------------------------
    @staticmethod
    def evaluate_sed_flux_iso_mono(nu, z, d_L, delta_D, mu_s, R_b, epsilon_0, u_0, n_e, *args,
                                   integrator=np_trapz, gamma=gamma_e_to_integrate,
                                   mu=mu_to_integrate, phi=phi_to_integrate):
        """EC on a monochromatic isotropic field (e.g. CMB), external_compton.py:63-142"""
        mu, phi = as_grid(mu, "mu"), as_grid(phi, "phi")
        tgt = [float(epsilon_0), strip(u_0, "erg cm-3", "u_0")]
        return _ec_call(_lib.EC_ISO_MONO, nu, z, d_L, delta_D, mu_s, R_b, tgt, n_e, args,
                        integrator, gamma, mu, mu.size, phi)

    @staticmethod
    def evaluate_sed_flux_ps_behind_jet(nu, z, d_L, delta_D, mu_s, R_b, epsilon_0, L_0, r, n_e,
                                        *args, integrator=np_trapz,
                                        gamma=gamma_e_to_integrate):
        """EC on a point source behind the jet, external_compton.py:163-240"""
        tgt = [float(epsilon_0), strip(L_0, "erg s-1", "L_0"), strip(r, "cm", "r")]
        return _ec_call(_lib.EC_PS_BEHIND_JET, nu, z, d_L, delta_D, mu_s, R_b, tgt, n_e, args,
                        integrator, gamma, None, 0, None)

    @staticmethod
    def evaluate_sed_flux_ss_disk(nu, z, d_L, delta_D, mu_s, R_b, M_BH, L_disk, eta, R_in, R_out,
                                  r, n_e, *args, integrator=np_trapz,
                                  gamma=gamma_e_to_integrate, mu_size=100, phi=phi_to_integrate):
        """EC on a Shakura-Sunyaev disk, external_compton.py:261-373"""
        tgt = [strip(M_BH, "g", "M_BH"), strip(L_disk, "erg s-1", "L_disk"), float(eta),
               strip(R_in, "cm", "R_in"), strip(R_out, "cm", "R_out"), strip(r, "cm", "r")]
        return _ec_call(_lib.EC_SS_DISK, nu, z, d_L, delta_D, mu_s, R_b, tgt, n_e, args,
                        integrator, gamma, None, int(mu_size), as_grid(phi, "phi"))

    @staticmethod
    def evaluate_sed_flux_blr(nu, z, d_L, delta_D, mu_s, R_b, L_disk, xi_line, epsilon_line,
                              R_line, r, n_e, *args, integrator=np_trapz,
                              gamma=gamma_e_to_integrate, mu=mu_to_integrate,
                              phi=phi_to_integrate):
        """EC on a spherical-shell BLR, external_compton.py:398-490"""
        mu, phi = as_grid(mu, "mu"), as_grid(phi, "phi")
        tgt = [strip(L_disk, "erg s-1", "L_disk"), float(xi_line), float(epsilon_line),
               strip(R_line, "cm", "R_line"), strip(r, "cm", "r")]
        return _ec_call(_lib.EC_BLR, nu, z, d_L, delta_D, mu_s, R_b, tgt, n_e, args, integrator,
                        gamma, mu, mu.size, phi)

    @staticmethod
    def evaluate_sed_flux_dt(nu, z, d_L, delta_D, mu_s, R_b, L_disk, xi_dt, epsilon_dt, R_dt, r,
                             n_e, *args, integrator=np_trapz, gamma=gamma_e_to_integrate,
                             phi=phi_to_integrate):
        """EC on a ring dust torus, external_compton.py:514-600"""
        tgt = [strip(L_disk, "erg s-1", "L_disk"), float(xi_dt), float(epsilon_dt),
               strip(R_dt, "cm", "R_dt"), strip(r, "cm", "r")]
        return _ec_call(_lib.EC_DT, nu, z, d_L, delta_D, mu_s, R_b, tgt, n_e, args, integrator,
                        gamma, None, 0, as_grid(phi, "phi"))

    # ---- instance conveniences ----------------------------------------------------------------
    def _blob_args(self):
        b = self.blob
        return b.z, b.d_L, b.delta_D, b.mu_s, b.R_b

    def sed_flux_cmb(self, nu):
        """external_compton.py:144-161"""
        b, t = self.blob, self.target
        return self.evaluate_sed_flux_iso_mono(
            nu, *self._blob_args(), t.epsilon_0, t.u_0, b.n_e, *b.n_e.parameters,
            integrator=self.integrator, gamma=b.gamma_e_external_frame, mu=self.mu, phi=self.phi)

    def sed_flux_ps_behind_jet(self, nu):
        """external_compton.py:242-259"""
        b, t = self.blob, self.target
        return self.evaluate_sed_flux_ps_behind_jet(
            nu, *self._blob_args(), t.epsilon_0, t.L_0, self.r, b.n_e, *b.n_e.parameters,
            integrator=self.integrator, gamma=b.gamma_e_external_frame)

    def sed_flux_ss_disk(self, nu):
        """external_compton.py:375-396"""
        b, t = self.blob, self.target
        return self.evaluate_sed_flux_ss_disk(
            nu, *self._blob_args(), t.M_BH, t.L_disk, t.eta, t.R_in, t.R_out, self.r, b.n_e,
            *b.n_e.parameters, integrator=self.integrator, gamma=b.gamma_e_external_frame,
            mu_size=self.mu_size, phi=self.phi)

    def sed_flux_blr(self, nu):
        """external_compton.py:492-512"""
        b, t = self.blob, self.target
        return self.evaluate_sed_flux_blr(
            nu, *self._blob_args(), t.L_disk, t.xi_line, t.epsilon_line, t.R_line, self.r, b.n_e,
            *b.n_e.parameters, integrator=self.integrator, gamma=b.gamma_e_external_frame,
            mu=self.mu, phi=self.phi)

    def sed_flux_dt(self, nu):
        """external_compton.py:602-621"""
        b, t = self.blob, self.target
        return self.evaluate_sed_flux_dt(
            nu, *self._blob_args(), t.L_disk, t.xi_dt, t.epsilon_dt, t.R_dt, self.r, b.n_e,
            *b.n_e.parameters, integrator=self.integrator, gamma=b.gamma_e_external_frame,
            phi=self.phi)

    def sed_flux(self, nu):
        """external_compton.py:623-634"""
        kind = _target_kind(self.target)
        fn = {"CMB": self.sed_flux_cmb, "PointSourceBehindJet": self.sed_flux_ps_behind_jet,
              "SSDisk": self.sed_flux_ss_disk, "SphericalShellBLR": self.sed_flux_blr,
              "RingDustTorus": self.sed_flux_dt}.get(kind)
        if fn is None:
            raise TypeError(f"unsupported target {type(self.target).__name__!r}")
        return fn(nu)

    def sed_luminosity(self, nu):
        d_L = strip(self.blob.d_L, "cm", "d_L")
        return wrap(4 * np.pi * d_L**2 * strip(self.sed_flux(nu), "erg cm-2 s-1"), "erg s-1")

    def sed_peak_flux(self, nu):
        return self.sed_flux(nu).max()

    def sed_peak_nu(self, nu):
        return nu[self.sed_flux(nu).argmax()]
