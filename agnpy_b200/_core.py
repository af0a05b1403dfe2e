"""Marshalling between numpy (CGS doubles) and the C ABI.  One function per entry point of
include/agnpy_b200.h, host-pointer flavour: used by the drop-in static methods.  The
device-pointer flavour is used by agnpy_b200.batch."""
import numpy as np

from . import _lib
from ._lib import check, f64, hptr, lib


def model_row(models, i, *, z=0.0, d_L=1.0, delta_D=1.0, mu_s=1.0, B=0.0, R_b=1.0, ne=None,
              tgt=()):
    m = models[i]
    m["z"], m["d_L"], m["delta_D"], m["mu_s"], m["B"], m["R_b"] = z, d_L, delta_D, mu_s, B, R_b
    if ne is not None:
        m["ne"][:] = ne
    t = np.zeros(8)
    t[:len(tgt)] = tgt
    m["tgt"][:] = t


def synch_sed_host(models, ne_kind, nu, gamma, ne_table=None, ssa_table=None, ssa=False,
                   integ=0, want_tau=False):
    nu, gamma = f64(nu), f64(gamma)
    n = len(models)
    sed = np.empty((n, nu.size))
    tau = np.empty((n, nu.size)) if want_tau else None
    check(lib().agnpy_b200_synch_sed_host(
        n, hptr(models), ne_kind, hptr(nu), nu.size, hptr(gamma), gamma.size,
        hptr(ne_table), hptr(ssa_table), int(bool(ssa)), integ, hptr(sed), hptr(tau)),
        "agnpy_b200_synch_sed_host")
    return (sed, tau) if want_tau else sed


def ssc_sed_host(models, ne_kind, nu, nu_seed, gamma, ne_table=None, ssa_table=None,
                 ssa=False, integ=0):
    nu, nu_seed, gamma = f64(nu), f64(nu_seed), f64(gamma)
    n = len(models)
    sed = np.empty((n, nu.size))
    check(lib().agnpy_b200_ssc_sed_host(
        n, hptr(models), ne_kind, hptr(nu), nu.size, hptr(nu_seed), nu_seed.size,
        hptr(gamma), gamma.size, hptr(ne_table), hptr(ssa_table), int(bool(ssa)), integ,
        hptr(sed)), "agnpy_b200_ssc_sed_host")
    return sed


def ec_sed_host(variant, models, ne_kind, nu, gamma, mu, n_mu, phi, ne_table=None, integ=0):
    nu, gamma = f64(nu), f64(gamma)
    mu = None if mu is None else f64(mu)
    phi = None if phi is None else f64(phi)
    n = len(models)
    sed = np.empty((n, nu.size))
    check(lib().agnpy_b200_ec_sed_host(
        variant, n, hptr(models), ne_kind, hptr(nu), nu.size, hptr(gamma), gamma.size,
        hptr(mu), int(n_mu), hptr(phi), 0 if phi is None else phi.size, hptr(ne_table),
        integ, hptr(sed)), "agnpy_b200_ec_sed_host")
    return sed


def tau_host(variant, models, nu, mu, n_a, phi, l_unit):
    nu = f64(nu)
    phi = None if phi is None else f64(phi)
    l_unit = None if l_unit is None else f64(l_unit)
    mu = None if mu is None else f64(mu)
    n = len(models)
    tau = np.empty((n, nu.size))
    check(lib().agnpy_b200_tau_host(
        variant, n, hptr(models), hptr(nu), nu.size, hptr(mu), int(n_a), hptr(phi),
        0 if phi is None else phi.size, hptr(l_unit), 0 if l_unit is None else l_unit.size,
        hptr(tau)), "agnpy_b200_tau_host")
    return tau


def bb_sed_host(variant, models, nu, nu_norm=None, n_R=100):
    nu = f64(nu)
    nu_norm = None if nu_norm is None else f64(nu_norm)
    n = len(models)
    sed = np.empty((n, nu.size))
    check(lib().agnpy_b200_bb_sed_host(
        variant, n, hptr(models), hptr(nu), nu.size, hptr(nu_norm),
        0 if nu_norm is None else nu_norm.size, int(n_R), hptr(sed)), "agnpy_b200_bb_sed_host")
    return sed


def tau_synch_host(models, ne_kind, nu, gamma, ne_table=None, ssa_table=None, n_s=200,
                   margin_low=1.0e-2, integ=0):
    nu, gamma = f64(nu), f64(gamma)
    n = len(models)
    tau = np.empty((n, nu.size))
    check(lib().agnpy_b200_tau_synch_host(
        n, hptr(models), ne_kind, hptr(nu), nu.size, hptr(gamma), gamma.size, hptr(ne_table),
        hptr(ssa_table), int(n_s), float(margin_low), integ, hptr(tau)), "agnpy_b200_tau_synch_host")
    return tau


def proton_synch_sed_host(models, ne_kind, nu, gamma, ne_table=None, integ=0):
    nu, gamma = f64(nu), f64(gamma)
    n = len(models)
    sed = np.empty((n, nu.size))
    check(lib().agnpy_b200_proton_synch_sed_host(
        n, hptr(models), ne_kind, hptr(nu), nu.size, hptr(gamma), gamma.size, hptr(ne_table), integ,
        hptr(sed)), "agnpy_b200_proton_synch_sed_host")
    return sed


def ssc_loss_rate_host(models, ne_kind, gamma_eval, nu_seed, gamma, ne_table=None):
    gamma_eval, nu_seed, gamma = f64(gamma_eval), f64(nu_seed), f64(gamma)
    n = len(models)
    out = np.empty((n, gamma_eval.size))
    check(lib().agnpy_b200_ssc_loss_rate_host(
        n, hptr(models), ne_kind, hptr(gamma_eval), gamma_eval.size, hptr(nu_seed), nu_seed.size,
        hptr(gamma), gamma.size, hptr(ne_table), hptr(out)), "agnpy_b200_ssc_loss_rate_host")
    return out


def target_u_host(variant, models, r, Gamma=None, n_mu=50):
    r = f64(np.ravel(r))
    n = len(models)
    Gamma = None if Gamma is None else f64(np.broadcast_to(Gamma, (n,)).copy())
    out = np.empty((n, r.size))
    check(lib().agnpy_b200_target_u_host(variant, n, hptr(models), hptr(Gamma), hptr(r), r.size,
                                         int(n_mu), hptr(out)), "agnpy_b200_target_u_host")
    return out


def ebl_absorption_host(z, nu, energy_ref, z_ref, values, sed=None):
    """absorption [n_models][n_nu]; with `sed` ([n_models][n_nu], C-contiguous float64) the
    table is applied in place and `sed` is returned"""
    z, nu = f64(np.atleast_1d(z)), f64(np.ravel(nu))
    energy_ref, z_ref, values = f64(energy_ref), f64(z_ref), f64(values)
    if values.shape != (energy_ref.size, z_ref.size):
        raise ValueError("EBL table must be values[n_energy][n_z]")
    if sed is None:
        out, mult = np.empty((z.size, nu.size)), 0
    else:
        if sed.shape != (z.size, nu.size) or sed.dtype != np.float64 or not sed.flags.c_contiguous:
            raise ValueError("sed must be a C-contiguous float64 array [n_models][n_nu]")
        out, mult = sed, 1
    check(lib().agnpy_b200_ebl_absorption_host(z.size, hptr(z), hptr(nu), nu.size, hptr(energy_ref),
                                               energy_ref.size, hptr(z_ref), z_ref.size, hptr(values),
                                               mult, hptr(out)), "agnpy_b200_ebl_absorption_host")
    return out


def trapz_loglog_host(rows, x, intervals=False):
    """rows [n_rows][n_x] integrated along the last axis with the reference's trapz_loglog"""
    rows, x = f64(rows), f64(x)
    n_rows, n_x = rows.shape
    out = np.empty((n_rows, n_x - 1) if intervals else (n_rows,))
    check(lib().agnpy_b200_trapz_loglog_host(hptr(rows), hptr(x), n_rows, n_x, int(bool(intervals)),
                                             hptr(out)), "agnpy_b200_trapz_loglog_host")
    return out
