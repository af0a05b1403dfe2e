"""Drop-in for agnpy/absorption/targets_absorption.py: `Absorption` with the reference's
static `evaluate_tau_{ss_disk,blr,dt}` signatures (targets_absorption.py:182-196, 284-297,
474-486), running on the fused FP64 CUDA kernel K4."""
import numpy as np

from . import _core, _lib, spectra
from .compton import _target_kind
from .grids import as_grid, mu_to_integrate, phi_to_integrate
from .units import strip

__all__ = ["Absorption"]


def _tau_call(variant, nu, z, mu_s, tgt, mu, n_a, phi, l_size, off_axis=False):
    nu_hz = strip(nu, "Hz", "nu")
    models = _lib.make_models(1)
    _core.model_row(models, 0, z=float(z), mu_s=float(mu_s), tgt=tgt)
    if l_size is None:
        l_unit = None
    elif off_axis:  # path measured from the emission point, targets_absorption.py:153, 405, 576
        l_unit = np.logspace(-5, 5, int(l_size))
    else:           # path measured from the black hole, targets_absorption.py:237, 333, 520
        l_unit = np.logspace(0, 5, int(l_size))
    phi = None if phi is None else as_grid(phi, "phi")
    tau = _core.tau_host(variant, models, np.ravel(nu_hz), mu, n_a, phi, l_unit)
    return tau[0].reshape(np.shape(nu_hz))


class Absorption:
    """targets_absorption.py:45-747; every branch of the reference's `tau()` dispatch."""

    def __init__(self, target, r=None, z=0, mu_s=1):
        self.target = target
        self.r = r
        self.z = z
        self.mu_s = mu_s
        self.set_mu()
        self.set_phi()
        self.set_l()
        if r is None and _target_kind(target) != "Blob" and not hasattr(target, "n_e"):
            raise ValueError(
                "No distance provided for absorption on " + str(target.__class__)
                + ", this can be only done for Blob class")

    def set_mu(self, mu_size=100):
        self.mu_size = mu_size
        self.mu = np.linspace(-1, 1, self.mu_size)

    def set_phi(self, phi_size=50):
        self.phi_size = phi_size
        self.phi = np.linspace(0, 2 * np.pi, self.phi_size)

    def set_l(self, l_size=50):
        self.l_size = l_size

    @staticmethod
    def evaluate_tau_ss_disk(nu, z, mu_s, M_BH, L_disk, eta, R_in, R_out, r, R_tilde_size=100,
                             l_tilde_size=50, phi=phi_to_integrate):
        """gamma-gamma opacity on a Shakura-Sunyaev disk, targets_absorption.py:182-264"""
        tgt = [strip(M_BH, "g", "M_BH"), strip(L_disk, "erg s-1", "L_disk"), float(eta),
               strip(R_in, "cm", "R_in"), strip(R_out, "cm", "R_out"), strip(r, "cm", "r")]
        return _tau_call(_lib.TAU_SS_DISK, nu, z, mu_s, tgt, None, int(R_tilde_size), phi,
                         l_tilde_size)

    @staticmethod
    def evaluate_tau_blr(nu, z, mu_s, L_disk, xi_line, epsilon_line, R_line, r, l_size=50,
                         mu=mu_to_integrate, phi=phi_to_integrate):
        """gamma-gamma opacity on a spherical-shell BLR, targets_absorption.py:284-353"""
        mu = as_grid(mu, "mu")
        tgt = [strip(L_disk, "erg s-1", "L_disk"), float(xi_line), float(epsilon_line),
               strip(R_line, "cm", "R_line"), strip(r, "cm", "r")]
        return _tau_call(_lib.TAU_BLR, nu, z, mu_s, tgt, mu, mu.size, phi, l_size)

    @staticmethod
    def evaluate_tau_dt(nu, z, mu_s, L_disk, xi_dt, epsilon_dt, R_dt, r, l_size=50,
                        phi=phi_to_integrate):
        """gamma-gamma opacity on a ring dust torus, targets_absorption.py:474-532"""
        tgt = [strip(L_disk, "erg s-1", "L_disk"), float(xi_dt), float(epsilon_dt),
               strip(R_dt, "cm", "R_dt"), strip(r, "cm", "r")]
        return _tau_call(_lib.TAU_DT, nu, z, mu_s, tgt, None, 0, phi, l_size)

    @staticmethod
    def evaluate_tau_ps_behind_blob(nu, z, mu_s, epsilon_0, L_0, r):
        """opacity on a point source behind the blob (closed form), targets_absorption.py:87-119"""
        tgt = [float(epsilon_0), strip(L_0, "erg s-1", "L_0"), strip(r, "cm", "r")]
        return _tau_call(_lib.TAU_PS_BEHIND_BLOB, nu, z, mu_s, tgt, None, 0, None, None)

    @staticmethod
    def evaluate_tau_ps_behind_blob_mu_s(nu, z, mu_s, epsilon_0, L_0, r, u_size=100):
        """same, integrated along an off-axis path, targets_absorption.py:128-173"""
        tgt = [float(epsilon_0), strip(L_0, "erg s-1", "L_0"), strip(r, "cm", "r")]
        return _tau_call(_lib.TAU_PS_BEHIND_BLOB_MU_S, nu, z, mu_s, tgt, None, 0, None, u_size,
                         off_axis=True)

    @staticmethod
    def evaluate_tau_blr_mu_s(nu, z, mu_s, L_disk, xi_line, epsilon_line, R_line, r, u_size=100,
                              mu=mu_to_integrate, phi=phi_to_integrate):
        """opacity on a spherical-shell BLR for any viewing angle, targets_absorption.py:355-436"""
        mu = as_grid(mu, "mu")
        tgt = [strip(L_disk, "erg s-1", "L_disk"), float(xi_line), float(epsilon_line),
               strip(R_line, "cm", "R_line"), strip(r, "cm", "r")]
        return _tau_call(_lib.TAU_BLR_MU_S, nu, z, mu_s, tgt, mu, mu.size, phi, u_size, off_axis=True)

    @staticmethod
    def evaluate_tau_dt_mu_s(nu, z, mu_s, L_disk, xi_dt, epsilon_dt, R_dt, r, u_size=100,
                             phi_re=phi_to_integrate):
        """opacity on a ring dust torus for any viewing angle, targets_absorption.py:534-597"""
        tgt = [strip(L_disk, "erg s-1", "L_disk"), float(xi_dt), float(epsilon_dt),
               strip(R_dt, "cm", "R_dt"), strip(r, "cm", "r")]
        return _tau_call(_lib.TAU_DT_MU_S, nu, z, mu_s, tgt, None, 0, phi_re, u_size, off_axis=True)

    def tau_ps_behind_blob(self, nu):
        """targets_absorption.py:121-126"""
        t = self.target
        return self.evaluate_tau_ps_behind_blob(nu, self.z, self.mu_s, t.epsilon_0, t.L_0, self.r)

    def tau_ps_behind_blob_mu_s(self, nu):
        """targets_absorption.py:175-180"""
        t = self.target
        return self.evaluate_tau_ps_behind_blob_mu_s(nu, self.z, self.mu_s, t.epsilon_0, t.L_0, self.r)

    def tau_blr_mu_s(self, nu):
        """targets_absorption.py:456-472"""
        t = self.target
        return self.evaluate_tau_blr_mu_s(nu, self.z, self.mu_s, t.L_disk, t.xi_line, t.epsilon_line,
                                          t.R_line, self.r, u_size=2 * self.l_size, mu=self.mu,
                                          phi=self.phi)

    def tau_dt_mu_s(self, nu):
        """targets_absorption.py:616-630"""
        t = self.target
        return self.evaluate_tau_dt_mu_s(nu, self.z, self.mu_s, t.L_disk, t.xi_dt, t.epsilon_dt,
                                         t.R_dt, self.r, u_size=2 * self.l_size, phi_re=self.phi)

    def tau_on_synchrotron(self, blob, nu, nu_s_size=200, delta_margin_low=1.0e-2):
        """opacity on the blob's own (self-absorbed) synchrotron photons,
        targets_absorption.py:629-694"""
        nu_hz = strip(nu, "Hz", "nu")
        n_e = blob.n_e
        kind, par = spectra.describe(n_e, n_e.parameters)
        gamma = as_grid(blob.gamma_e, "gamma")
        ne_t = ssa_t = None
        if kind == _lib.NE_TABLE:
            g = spectra.snap_boundaries(gamma, par[1], par[2])
            ne_t, ssa_t = spectra.tables(n_e, n_e.parameters, g, True)
        models = _lib.make_models(1)
        _core.model_row(models, 0, z=float(blob.z), d_L=strip(blob.d_L, "cm", "d_L"),
                        delta_D=float(blob.delta_D), B=strip(blob.B, "G", "B"),
                        R_b=strip(blob.R_b, "cm", "R_b"), ne=par)
        tau = _core.tau_synch_host(models, kind, np.ravel(nu_hz), gamma, ne_t, ssa_t, nu_s_size,
                                   delta_margin_low)
        return tau[0].reshape(np.shape(nu_hz))

    def tau_ss_disk(self, nu):
        """targets_absorption.py:266-282"""
        t = self.target
        return self.evaluate_tau_ss_disk(nu, self.z, self.mu_s, t.M_BH, t.L_disk, t.eta, t.R_in,
                                         t.R_out, self.r, R_tilde_size=100,
                                         l_tilde_size=self.l_size, phi=self.phi)

    def tau_blr(self, nu):
        """targets_absorption.py:438-454"""
        t = self.target
        return self.evaluate_tau_blr(nu, self.z, self.mu_s, t.L_disk, t.xi_line, t.epsilon_line,
                                     t.R_line, self.r, l_size=self.l_size, mu=self.mu,
                                     phi=self.phi)

    def tau_dt(self, nu):
        """targets_absorption.py:599-614"""
        t = self.target
        return self.evaluate_tau_dt(nu, self.z, self.mu_s, t.L_disk, t.xi_dt, t.epsilon_dt,
                                    t.R_dt, self.r, l_size=self.l_size, phi=self.phi)

    def tau(self, nu):
        """targets_absorption.py:696-730"""
        kind = _target_kind(self.target)
        if kind == "Blob" or hasattr(self.target, "n_e"):
            return self.tau_on_synchrotron(self.target, nu)
        if self.mu_s == 1:
            table = {"PointSourceBehindJet": self.tau_ps_behind_blob, "SSDisk": self.tau_ss_disk,
                     "SphericalShellBLR": self.tau_blr, "RingDustTorus": self.tau_dt}
        else:  # the reference has no off-axis disk opacity either (:724-726)
            table = {"PointSourceBehindJet": self.tau_ps_behind_blob_mu_s,
                     "SphericalShellBLR": self.tau_blr_mu_s, "RingDustTorus": self.tau_dt_mu_s}
            if kind == "SSDisk":
                raise NotImplementedError("off-axis opacity on the disk does not exist in agnpy")
        fn = table.get(kind)
        if fn is None:
            raise TypeError(f"unsupported target {type(self.target).__name__!r}")
        return fn(nu)

    def absorption(self, nu):
        """targets_absorption.py:732-736"""
        return np.exp(-self.tau(nu))

    def absorption_homogeneous(self, nu):
        """targets_absorption.py:738-747"""
        t = self.tau(nu)
        return (1 - np.exp(-t)) / t
