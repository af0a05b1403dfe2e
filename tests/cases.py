"""Physical test cases shared by the oracle tests and the GPU parity tests.

All quantities are CGS doubles.  Inputs follow SURVEY.md 8(d) "Synthetic inputs" and the
reference's own test modules (file:line cited per case).
"""
import numpy as np

from oracle import agnpy_oracle as orc

GOLDEN = __import__("pathlib").Path(__file__).resolve().parent / "golden"


# --------------------------------------------------------------------------
# tiny cosmology helper for the literature cases (the kernels take d_L as an input; the
# reference obtains it from astropy's Planck18, emission_regions/blob.py:81)
# --------------------------------------------------------------------------
def lum_dist_cm(z, H0=67.66, Om0=0.30966, Tcmb=2.7255, Neff=3.046):
    """flat LCDM with photons + massless neutrinos; agrees with astropy Planck18 to <1e-3
    (SURVEY.md 8(c)), irrelevant at the 5-35 % tolerances of the literature curves."""
    h = H0 / 100.0
    Ogam = 2.472e-5 * (Tcmb / 2.7255) ** 4 / h**2
    Orad = Ogam * (1 + 0.2271 * Neff)
    Ol = 1 - Om0 - Orad
    zz = np.linspace(0, z, 20001)
    E = np.sqrt(Om0 * (1 + zz) ** 3 + Orad * (1 + zz) ** 4 + Ol)
    dc = np.trapezoid(1 / E, zz) * (orc.C_LIGHT / 1e5 / H0)  # Mpc
    return (1 + z) * dc * 3.0856775814913673e24


def z_from_lum_dist(d_L_cm):
    lo, hi = 0.0, 10.0
    for _ in range(80):
        mid = 0.5 * (lo + hi)
        if lum_dist_cm(mid) < d_L_cm:
            lo = mid
        else:
            hi = mid
    return 0.5 * (lo + hi)


def load_curve(rel):
    """agnpy/utils/validation_utils.py:51-56 (`extract_columns_sample_file`)"""
    t = np.loadtxt(GOLDEN / "reference_data" / rel, delimiter=",", comments="#")
    return t[:, 0], t[:, 1]


def within(x, y_comp, y_ref, rtol, x_range=None):
    """agnpy/utils/validation_utils.py:59-72 (`check_deviation`)"""
    if x_range is not None:
        sel = (x >= x_range[0]) * (x <= x_range[1])
        y_comp, y_ref = y_comp[sel], y_ref[sel]
    return np.allclose(y_comp, y_ref, atol=0, rtol=rtol)


def max_rel_dev(a, b):
    """parity metric of SURVEY.md 8(d): max |a/b - 1| where b is not negligible; zeros
    must agree exactly."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    scale = np.max(np.abs(b)) if b.size else 0.0
    big = np.abs(b) > 1e-300 * max(scale, 1e-300)
    big &= np.abs(b) > 0
    dev = 0.0 if not big.any() else float(np.max(np.abs(a[big] / b[big] - 1)))
    zeros_ok = bool(np.all(a[~big] == 0) or np.all(np.abs(a[~big]) <= 1e-300 * scale))
    return dev, zeros_ok


# --------------------------------------------------------------------------
# C1: tutorial blob (docs/snippets/blob_snippet.py:16-26)
# --------------------------------------------------------------------------
def case_c1(n_nu=100):
    ne = ("BrokenPowerLaw", (1e-8, 1.9, 2.6, 1e4, 10.0, 1e6))
    return dict(
        nu=np.logspace(8, 28, n_nu), z=0.069, d_L=1e27, delta_D=10.0, B=1.0, R_b=1e16,
        ne_kind=ne[0], ne_args=ne[1], gamma=np.logspace(1, 6, 200),
    )


# --------------------------------------------------------------------------
# C2/C3: Finke 2016 EC blob (compton/tests/test_compton.py:41-66)
# --------------------------------------------------------------------------
R_B_EC = 1e16
V_B_EC = 4 / 3 * np.pi * R_B_EC**3
Z_EC = 1.0
D_L_EC = 2.0957e28  # fixed synthetic input ~ Planck18 d_L(z=1)  
DELTA_EC = 40.0
GAMMA_BULK_EC = 40.0
BETA_EC = np.sqrt(1 - 1 / GAMMA_BULK_EC**2)
MU_S_EC = (1 - 1 / (GAMMA_BULK_EC * DELTA_EC)) / BETA_EC  # emission_regions/blob.py:109-112
L_DISK = 2e46
M_BH = 1.2e9 * orc.M_SUN
ETA = 1 / 12
R_G = orc.G_NEWTON * M_BH / orc.C_LIGHT**2
XI_LINE, R_LINE = 0.024, 1e17
EPS_LINE = orc.line_epsilon()
XI_DT, T_DT = 0.1, 1e3
EPS_DT = orc.dt_epsilon(T_DT)
R_DT = orc.dt_sublimation_radius(L_DISK, T_DT)


def ne_ec():
    shape = (2.0, 3.5, 1e4, 20.0, 5e7)
    k = orc.ne_norm_from_total_energy("BrokenPowerLaw", 6e42, V_B_EC, *shape)
    return "BrokenPowerLaw", (k,) + shape


def blob_ec():
    kind, args = ne_ec()
    return dict(z=Z_EC, d_L=D_L_EC, delta_D=DELTA_EC, mu_s=MU_S_EC, R_b=R_B_EC,
                ne_kind=kind, ne_args=args)


def case_c2(n_nu=200, r=1e18, n_gamma=200, n_mu=100, n_phi=50):
    b = blob_ec()
    b.update(nu=np.logspace(15, 30, n_nu), L_disk=L_DISK, xi_line=XI_LINE,
             epsilon_line=EPS_LINE, R_line=R_LINE, r=r,
             gamma=np.logspace(1, 9, n_gamma), mu=np.linspace(-1, 1, n_mu),
             phi=np.linspace(0, 2 * np.pi, n_phi))
    return b


def case_c3_dt(n_nu=200, r=1e20, n_gamma=200, n_phi=50):
    b = blob_ec()
    b.update(nu=np.logspace(15, 30, n_nu), L_disk=L_DISK, xi_dt=XI_DT, epsilon_dt=EPS_DT,
             R_dt=R_DT, r=r, gamma=np.logspace(1, 9, n_gamma),
             phi=np.linspace(0, 2 * np.pi, n_phi))
    return b


def case_disk(n_nu=60, r=1e17, n_gamma=200, mu_size=100, n_phi=50):
    b = blob_ec()
    b.update(nu=np.logspace(15, 30, n_nu), M_BH=M_BH, L_disk=L_DISK, eta=ETA,
             R_in=6 * R_G, R_out=200 * R_G, r=r, gamma=np.logspace(1, 9, n_gamma),
             mu_size=mu_size, phi=np.linspace(0, 2 * np.pi, n_phi))
    return b


def case_tau(n_nu=200):
    """docs/snippets/absorption_snippet.py:13-29: E = logspace(0,5) GeV, z=0.859, r=1.1e16"""
    E_GeV = np.logspace(0, 5, n_nu)
    nu = E_GeV * 1.602176634e-3 / orc.H_PLANCK
    return dict(nu=nu, z=0.859, mu_s=1.0, r=1.1e16)


ORACLE_EC = {
    "blr": lambda c, integ="trapz": orc.ec_blr_sed_flux(
        c["nu"], c["z"], c["d_L"], c["delta_D"], c["mu_s"], c["R_b"], c["L_disk"],
        c["xi_line"], c["epsilon_line"], c["R_line"], c["r"], c["ne_kind"], c["ne_args"],
        integrator=integ, gamma=c["gamma"], mu=c["mu"], phi=c["phi"]),
    "dt": lambda c, integ="trapz": orc.ec_dt_sed_flux(
        c["nu"], c["z"], c["d_L"], c["delta_D"], c["mu_s"], c["R_b"], c["L_disk"],
        c["xi_dt"], c["epsilon_dt"], c["R_dt"], c["r"], c["ne_kind"], c["ne_args"],
        integrator=integ, gamma=c["gamma"], phi=c["phi"]),
    "ss_disk": lambda c, integ="trapz": orc.ec_ss_disk_sed_flux(
        c["nu"], c["z"], c["d_L"], c["delta_D"], c["mu_s"], c["R_b"], c["M_BH"], c["L_disk"],
        c["eta"], c["R_in"], c["R_out"], c["r"], c["ne_kind"], c["ne_args"],
        integrator=integ, gamma=c["gamma"], mu_size=c["mu_size"], phi=c["phi"]),
}


# --------------------------------------------------------------------------
# C5: fit parameter vectors in the sherpa wrapper's flat order (ranges inside the
# fit bounds of agnpy/fit/core.py:92-113, 151-157)
# --------------------------------------------------------------------------
C5_NU = np.logspace(10, 28, 100)
# column ranges of the 23-parameter "ec_blr_dt" vector that each sherpa scenario takes
# (agnpy/fit/sherpa_wrapper.py:55, 83-98, 157-172, 233-251)
C5_COLUMNS = {
    "ssc": list(range(0, 10)),
    "ec_blr": list(range(0, 20)),
    "ec_dt": list(range(0, 17)) + [20, 21, 22],
    "ec_blr_dt": list(range(0, 23)),
}


def c5_parameter_vectors(B, seed=20261019):
    """[B, 23]: log10_k, p1, p2, log10_gamma_b, log10_gamma_min, log10_gamma_max, z, delta_D, log10_B,
    t_var, mu_s, log10_r, log10_L_disk, M_BH, m_dot, R_in, R_out, xi_line, lambda_line, R_line, xi_dt,
    T_dt, R_dt"""
    rng = np.random.default_rng(seed)
    pars = np.zeros((B, 23))
    delta = rng.uniform(10, 40, B)
    pars[:, 0], pars[:, 1], pars[:, 2] = rng.uniform(-9, -7, B), rng.uniform(1.5, 2.5, B), rng.uniform(3, 4, B)
    pars[:, 3], pars[:, 4], pars[:, 5] = rng.uniform(3.5, 5, B), np.log10(20), np.log10(5e7)
    pars[:, 6], pars[:, 7], pars[:, 8], pars[:, 9] = 1.0, delta, rng.uniform(-1.5, 0, B), 86400.0
    pars[:, 10] = (1 - 1 / delta**2) / np.sqrt(1 - 1 / delta**2)
    pars[:, 11], pars[:, 12] = rng.uniform(17, 19, B), np.log10(2e46)
    pars[:, 13:] = [M_BH, 1e26, 6 * R_G, 200 * R_G, 0.024, 1215.67, 1e17, 0.1, 1e3, R_DT]
    return pars
