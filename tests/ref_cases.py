"""Cases on which the oracle (and, in the GPU tests, the CUDA path) is pinned against outputs of the
reference's OWN, unmodified `evaluate_*` code.

Each case has two sides on identical CGS inputs:
  * `ref(A)`  -- calls the reference (`A.agnpy`, Quantities built with `A.u`), run by
                 tests/golden/make_ref_evaluate.py in the build container -> tests/golden/ref_evaluate.npz;
  * `orc(idx)`-- the oracle restatement (`oracle/agnpy_oracle.py`), on the frequency subset `idx` so the
                 CPU suite stays within minutes (every output frequency is independent).
`tests/test_oracle_vs_reference.py` asserts oracle == reference at RTOL_ORACLE_VS_REF (the two differ only
by the rounding of SI->CGS conversion factors inside astropy-style unit arithmetic).

Inputs: SURVEY.md 8(d) synthetic inputs = BASELINE.json configs 1-3 (+ every other entry point).
"""
import numpy as np

from oracle import agnpy_oracle as orc
from tests import cases

RTOL_ORACLE_VS_REF = 1e-13

CASES = {}


# The two places where the reference's own formulas amplify last-bit input differences (the SI->CGS
# factor roundings) beyond 1e-13; the oracle evaluates the same formulas, so these are properties of the
# reference arithmetic, not restatement slack:
#  * exp(-x) in R(x) / Planck's law: relative error x * eps; the far tails (x ~ 700, SED 1e-250 of its
#    peak) reach 2.3e-13 (synchrotron.py:16-24);
#  * tau_to_attenuation (synchrotron.py:64-68): 1/2 + e^-tau/tau - (1-e^-tau)/tau^2 cancels to
#    ~tau/3 from terms of order 1/tau^2, i.e. it returns noise of relative size ~ eps/tau^3 just above
#    its tau = 1e-3 switch (measured: up to 8e-10).  tau itself is pinned at RTOL_EXP_TAIL.
RTOL_EXP_TAIL = 5e-13
RTOL_SSA_ATTENUATION = 2e-9


def case(name, sub=None, rtol=RTOL_ORACLE_VS_REF):
    """register `fn(side, A, idx[, store])`; `sub` = stride of the frequency subset the oracle side
    computes; `store` = the fixture dict, for auxiliary values a case records on the reference side
    (a d_L the reference derived from z) and reads back on the oracle side"""
    def deco(fn):
        CASES[name] = dict(fn=fn, sub=sub, rtol=rtol)
        return fn
    return deco


def run(name, side, A=None, store=None):
    """full output on the reference side; (subset indices, output on them) on the oracle side"""
    entry = CASES[name]
    fn = entry["fn"]
    kw = {"store": store} if "store" in fn.__code__.co_varnames[:fn.__code__.co_argcount] else {}
    if side == "ref":
        return fn("ref", A, None, **kw)
    if entry["sub"] in (None, 1):
        return None, fn("orc", None, None, **kw)
    n = len(store[name])
    idx = np.arange(0, n, entry["sub"])
    return idx, fn("orc", None, idx, **kw)


class RefSide:
    """what the reference side of a case needs: the imported reference and its units module"""

    def __init__(self, agnpy):
        import astropy.units as u
        import astropy.constants as const
        self.agnpy, self.u, self.const = agnpy, u, const

    def Q(self, value, unit):
        return np.asarray(value, dtype=float) * self.u.Unit(unit)

    def ne(self, kind):
        return getattr(self.agnpy, kind)

    def ne_args(self, kind, args):
        """k carries cm-3 (spectra.py `parameters`), the shape parameters are plain"""
        return (self.Q(args[0], "cm-3"),) + tuple(args[1:])

    def integ(self, name):
        import importlib
        ref_math = importlib.import_module("agnpy.utils.math")
        return {"trapz": np.trapz, "trapz_loglog": ref_math.trapz_loglog}[name]


class ProductSide(RefSide):
    """the SAME case code as the reference side, run on `agnpy_b200` (the drop-in): class names,
    argument order, Quantities and keyword arguments are those of the reference calls"""
    is_product = True

    def __init__(self):
        import agnpy_b200
        self.agnpy, self.u, self.const = agnpy_b200, agnpy_b200.u, None

    def integ(self, name):
        from agnpy_b200 import grids
        return {"trapz": grids.np_trapz, "trapz_loglog": self.agnpy.trapz_loglog}[name]

    def to_R_g_units(self, r_cm, M_g):
        from agnpy_b200 import constants as const
        return r_cm / (const.G * M_g / const.c**2)

    def fit_sherpa(self, scenario, pars, nu):
        from agnpy_b200.fit import FitSedBatch
        model = FitSedBatch(nu, "BrokenPowerLaw", scenario=scenario)
        return model.sherpa(nu * orc.H_PLANCK / EV_ERG, pars).T

    def fit_gammapy(self, targets, pars, energy_eV):
        from agnpy_b200.fit import FitSedBatch, scenario_of
        model = FitSedBatch(energy_eV * EV_ERG / orc.H_PLANCK, "BrokenPowerLaw",
                            scenario=scenario_of(targets))
        kwargs = {n: pars[:, i] for i, n in enumerate(GAMMAPY_NAMES)}
        return model.gammapy(energy_eV, norm=1.0, **kwargs).T


def _val(x):
    """plain ndarray of a reference result (SEDs are Quantities in erg cm-2 s-1, taus plain)"""
    return np.asarray(getattr(x, "value", x), dtype=float)


def _chunked(fn, nu_hz, chunk):
    return np.concatenate([np.atleast_1d(_val(fn(nu_hz[i:i + chunk])))
                           for i in range(0, len(nu_hz), chunk)])


# ------------------------------------------------------------------------------------------
# config 1: Synchrotron + SSC of the tutorial blob (docs/snippets/blob_snippet.py:16-26)
# ------------------------------------------------------------------------------------------
def _c1(side, A, idx, what, integ, ssa):
    c = cases.case_c1()
    nu = c["nu"] if idx is None else c["nu"][idx]
    if side == "orc":
        fn = {"synch": orc.synch_sed_flux, "ssc": orc.ssc_sed_flux, "tau_ssa": orc.synch_tau_ssa}[what]
        kw = {} if what == "tau_ssa" else {"ssa": ssa}
        return fn(nu, c["z"], c["d_L"], c["delta_D"], c["B"], c["R_b"], c["ne_kind"], c["ne_args"],
                  integrator=integ, gamma=c["gamma"].copy(), **kw)
    cls = {"synch": A.agnpy.Synchrotron, "tau_ssa": A.agnpy.Synchrotron,
           "ssc": A.agnpy.SynchrotronSelfCompton}[what]
    fn = cls.evaluate_tau_ssa if what == "tau_ssa" else cls.evaluate_sed_flux
    kw = {} if what == "tau_ssa" else {"ssa": ssa}
    return _val(fn(A.Q(nu, "Hz"), c["z"], A.Q(c["d_L"], "cm"), c["delta_D"], A.Q(c["B"], "G"),
                   A.Q(c["R_b"], "cm"), A.ne(c["ne_kind"]), *A.ne_args(c["ne_kind"], c["ne_args"]),
                   integrator=A.integ(integ), gamma=c["gamma"].copy(), **kw))


for _what in ("synch", "ssc"):
    for _integ in ("trapz", "trapz_loglog"):
        for _ssa in (False, True):
            case(f"c1_{_what}_{_integ}" + ("_ssa" if _ssa else ""), sub=(1 if _what == "synch" else 4),
                 rtol=RTOL_SSA_ATTENUATION if _ssa else (RTOL_EXP_TAIL if _what == "synch" else
                                                         RTOL_ORACLE_VS_REF))(
                lambda side, A, idx, w=_what, i=_integ, s=_ssa: _c1(side, A, idx, w, i, s))
case("c1_tau_ssa_trapz", rtol=RTOL_EXP_TAIL)(lambda side, A, idx: _c1(side, A, idx, "tau_ssa", "trapz", False))
case("c1_tau_ssa_trapz_loglog", rtol=RTOL_EXP_TAIL)(lambda side, A, idx: _c1(side, A, idx, "tau_ssa", "trapz_loglog", False))

# every electron-distribution shape through Synchrotron (n_e evaluate + SSA integrand, a21)
SHAPES = {
    "PowerLaw": (1e-6, 2.3, 10.0, 1e6),
    "BrokenPowerLaw": (1e-8, 1.9, 2.6, 1e4, 10.0, 1e6),
    "LogParabola": (1e-7, 2.1, 0.3, 1e3, 10.0, 1e6),
    "ExpCutoffPowerLaw": (1e-6, 2.0, 3e4, 10.0, 1e6),
    "ExpCutoffBrokenPowerLaw": (1e-7, 1.8, 2.9, 2e5, 5e3, 10.0, 1e6),
}


def _shape(side, A, idx, kind, ssa):
    c = cases.case_c1(n_nu=40)
    if side == "orc":
        return orc.synch_sed_flux(c["nu"], c["z"], c["d_L"], c["delta_D"], c["B"], c["R_b"], kind,
                                  SHAPES[kind], ssa=ssa, integrator="trapz", gamma=c["gamma"].copy())
    return _val(A.agnpy.Synchrotron.evaluate_sed_flux(
        A.Q(c["nu"], "Hz"), c["z"], A.Q(c["d_L"], "cm"), c["delta_D"], A.Q(c["B"], "G"),
        A.Q(c["R_b"], "cm"), A.ne(kind), *A.ne_args(kind, SHAPES[kind]), ssa=ssa,
        integrator=np.trapz, gamma=c["gamma"].copy()))


for _kind in SHAPES:
    for _ssa in (False, True):
        case(f"shape_synch_{_kind}" + ("_ssa" if _ssa else ""),
             rtol=RTOL_SSA_ATTENUATION if _ssa else RTOL_EXP_TAIL)(
            lambda side, A, idx, k=_kind, s=_ssa: _shape(side, A, idx, k, s))


@case("c1_proton_synch", rtol=RTOL_EXP_TAIL)
def _proton(side, A, idx):
    c = cases.case_c1(n_nu=40)
    args = (1e-3, 2.2, 10.0, 1e8)
    gamma = np.logspace(1, 8, 200)
    if side == "orc":
        return orc.proton_synch_sed_flux(c["nu"], c["z"], c["d_L"], c["delta_D"], 10.0, c["R_b"],
                                         "PowerLaw", args, gamma=gamma)
    return _val(A.agnpy.ProtonSynchrotron.evaluate_sed_flux(
        A.Q(c["nu"], "Hz"), c["z"], A.Q(c["d_L"], "cm"), c["delta_D"], A.Q(10.0, "G"),
        A.Q(c["R_b"], "cm"), A.ne("PowerLaw"), *A.ne_args("PowerLaw", args), gamma=gamma))


# ------------------------------------------------------------------------------------------
# configs 2/3: External Compton (compton/tests/test_compton.py:41-66 blob)
# ------------------------------------------------------------------------------------------
def _blob_q(A, c):
    return (c["z"], A.Q(c["d_L"], "cm"), c["delta_D"], c["mu_s"], A.Q(c["R_b"], "cm"))


def _ne_q(A, c):
    return (A.ne(c["ne_kind"]),) + A.ne_args(c["ne_kind"], c["ne_args"])


def _ec_blr(side, A, idx, r, integ, n_nu=200):
    c = cases.case_c2(n_nu=n_nu, r=r)
    nu = c["nu"] if idx is None else c["nu"][idx]
    if side == "orc":
        return orc.chunked_over_nu(
            lambda nn: cases.ORACLE_EC["blr"](dict(c, nu=nn), integ), nu, 10)
    EC = A.agnpy.ExternalCompton
    return _chunked(lambda nn: EC.evaluate_sed_flux_blr(
        A.Q(nn, "Hz"), *_blob_q(A, c), A.Q(c["L_disk"], "erg s-1"), c["xi_line"], c["epsilon_line"],
        A.Q(c["R_line"], "cm"), A.Q(c["r"], "cm"), *_ne_q(A, c), integrator=A.integ(integ),
        gamma=c["gamma"].copy(), mu=c["mu"], phi=c["phi"]), nu, 10)


for _r in (1e16, 2e17, 1e18):
    case(f"c2_ec_blr_r{_r:g}", sub=8)(lambda side, A, idx, r=_r: _ec_blr(side, A, idx, r, "trapz"))
case("c2_ec_blr_loglog_r1e18", sub=2)(
    lambda side, A, idx: _ec_blr(side, A, idx, 1e18, "trapz_loglog", n_nu=20))


def _ec_dt(side, A, idx, integ):
    c = cases.case_c3_dt()
    nu = c["nu"] if idx is None else c["nu"][idx]
    if side == "orc":
        return cases.ORACLE_EC["dt"](dict(c, nu=nu), integ)
    EC = A.agnpy.ExternalCompton
    return _val(EC.evaluate_sed_flux_dt(
        A.Q(nu, "Hz"), *_blob_q(A, c), A.Q(c["L_disk"], "erg s-1"), c["xi_dt"], c["epsilon_dt"],
        A.Q(c["R_dt"], "cm"), A.Q(c["r"], "cm"), *_ne_q(A, c), integrator=A.integ(integ),
        gamma=c["gamma"].copy(), phi=c["phi"]))


case("c3_ec_dt")(lambda side, A, idx: _ec_dt(side, A, idx, "trapz"))
case("c3_ec_dt_loglog")(lambda side, A, idx: _ec_dt(side, A, idx, "trapz_loglog"))


def _ec_disk(side, A, idx, r):
    c = cases.case_disk(n_nu=40, r=r)
    nu = c["nu"] if idx is None else c["nu"][idx]
    if side == "orc":
        return orc.chunked_over_nu(lambda nn: cases.ORACLE_EC["ss_disk"](dict(c, nu=nn)), nu, 10)
    EC = A.agnpy.ExternalCompton
    return _chunked(lambda nn: EC.evaluate_sed_flux_ss_disk(
        A.Q(nn, "Hz"), *_blob_q(A, c), A.Q(c["M_BH"], "g"), A.Q(c["L_disk"], "erg s-1"), c["eta"],
        A.Q(c["R_in"], "cm"), A.Q(c["R_out"], "cm"), A.Q(c["r"], "cm"), *_ne_q(A, c),
        integrator=np.trapz, gamma=c["gamma"].copy(), mu_size=c["mu_size"], phi=c["phi"]), nu, 10)


case("ec_ss_disk_r1e17", sub=4)(lambda side, A, idx: _ec_disk(side, A, idx, 1e17))
case("ec_ss_disk_r1e18", sub=4)(lambda side, A, idx: _ec_disk(side, A, idx, 1e18))

# CMB at z = 1 (targets/targets.py:164-169): u_0 = 4 sigma_sb/c T^4 (1+z)^4, eps_0 = 2.7 k_B T/mec2 (1+z)
EPS_CMB = 2.7 * orc.K_B * 2.72548 / orc.MEC2 * 2.0
U_CMB = 4 * 5.670374419e-5 / orc.C_LIGHT * 2.72548**4 * 16.0


@case("ec_iso_mono_cmb", sub=4)
def _ec_cmb(side, A, idx):
    c = cases.case_c2(n_nu=40)
    nu = c["nu"] if idx is None else c["nu"][idx]
    if side == "orc":
        return orc.chunked_over_nu(lambda nn: orc.ec_iso_mono_sed_flux(
            nn, c["z"], c["d_L"], c["delta_D"], c["mu_s"], c["R_b"], EPS_CMB, U_CMB, c["ne_kind"],
            c["ne_args"], gamma=c["gamma"].copy(), mu=c["mu"], phi=c["phi"]), nu, 10)
    EC = A.agnpy.ExternalCompton
    return _chunked(lambda nn: EC.evaluate_sed_flux_iso_mono(
        A.Q(nn, "Hz"), *_blob_q(A, c), EPS_CMB, A.Q(U_CMB, "erg cm-3"), *_ne_q(A, c),
        integrator=np.trapz, gamma=c["gamma"].copy(), mu=c["mu"], phi=c["phi"]), nu, 10)


@case("ec_ps_behind_jet")
def _ec_ps(side, A, idx):
    c = cases.case_c2(n_nu=60)
    if side == "orc":
        return orc.ec_ps_behind_jet_sed_flux(c["nu"], c["z"], c["d_L"], c["delta_D"], c["mu_s"],
                                             c["R_b"], 1e-5, 2e46, 1e18, c["ne_kind"], c["ne_args"],
                                             gamma=c["gamma"].copy())
    EC = A.agnpy.ExternalCompton
    return _val(EC.evaluate_sed_flux_ps_behind_jet(
        A.Q(c["nu"], "Hz"), *_blob_q(A, c), 1e-5, A.Q(2e46, "erg s-1"), A.Q(1e18, "cm"),
        *_ne_q(A, c), integrator=np.trapz, gamma=c["gamma"].copy()))


# ------------------------------------------------------------------------------------------
# random parameter sweep: every process of 8(a) on 8 random blobs / targets / electron spectra at small
# grids (the BASELINE cases above fix the physical parameters; this one moves all of them)
# ------------------------------------------------------------------------------------------
SWEEP_KINDS = ("PowerLaw", "BrokenPowerLaw", "LogParabola", "ExpCutoffPowerLaw", "ExpCutoffBrokenPowerLaw")


def _sweep_draw(seed):
    rng = np.random.default_rng(4200 + seed)
    kind = SWEEP_KINDS[seed % len(SWEEP_KINDS)]
    gmin, gmax = 10 ** rng.uniform(0.5, 2), 10 ** rng.uniform(5, 7.5)
    k = 10 ** rng.uniform(-9, -5)
    ne_args = {
        "PowerLaw": (k, rng.uniform(1.8, 3), gmin, gmax),
        "BrokenPowerLaw": (k, rng.uniform(1.5, 2.5), rng.uniform(2.6, 4), 10 ** rng.uniform(3, 4.5), gmin, gmax),
        "LogParabola": (k, rng.uniform(1.8, 2.6), rng.uniform(0.05, 0.6), 10 ** rng.uniform(2.5, 4), gmin, gmax),
        "ExpCutoffPowerLaw": (k, rng.uniform(1.5, 2.8), 10 ** rng.uniform(4, 6), gmin, gmax),
        "ExpCutoffBrokenPowerLaw": (k, rng.uniform(1.5, 2.4), rng.uniform(2.6, 3.8), 10 ** rng.uniform(4.5, 6),
                                    10 ** rng.uniform(3, 4), gmin, gmax),
    }[kind]
    M_BH = 10 ** rng.uniform(8, 9.5) * 1.988409870698051e33
    R_g = orc.G_NEWTON * M_BH / orc.C_LIGHT**2
    L_disk = 10 ** rng.uniform(45, 46.7)
    return dict(
        ne_kind=kind, ne_args=ne_args, z=rng.uniform(0.05, 2.5), d_L=10 ** rng.uniform(27, 28.7),
        delta_D=rng.uniform(3, 45), mu_s=rng.uniform(0.3, 0.99995), B=10 ** rng.uniform(-2, 1),
        R_b=10 ** rng.uniform(15, 17), L_disk=L_disk, xi_line=rng.uniform(0.01, 0.3),
        epsilon_line=cases.EPS_LINE * rng.uniform(0.5, 2), R_line=10 ** rng.uniform(16.5, 17.5),
        xi_dt=rng.uniform(0.05, 0.6), epsilon_dt=cases.EPS_DT * rng.uniform(0.5, 2),
        R_dt=10 ** rng.uniform(18, 19), r=10 ** rng.uniform(16, 19.5), M_BH=M_BH, eta=rng.uniform(0.05, 0.2),
        R_in=6 * R_g, R_out=rng.uniform(100, 500) * R_g,
        nu=np.logspace(rng.uniform(9, 12), rng.uniform(26, 29), 12), nu_tau=np.logspace(24, 29, 10),
        gamma=np.logspace(np.log10(gmin) - 0.3, np.log10(gmax) + 0.2, 70),
        gamma_ec=np.logspace(0.5, 8.5, 80), mu=np.linspace(-1, 1, 18), phi=np.linspace(0, 2 * np.pi, 14))


@case("random_sweep", rtol=RTOL_EXP_TAIL)
def _random_sweep(side, A, idx):
    out = []
    for seed in range(8):
        c = _sweep_draw(seed)
        blob = (c["z"], c["d_L"], c["delta_D"])
        if side == "orc":
            ne = (c["ne_kind"], c["ne_args"])
            out += [
                orc.synch_sed_flux(c["nu"], *blob, c["B"], c["R_b"], *ne, gamma=c["gamma"].copy()),
                orc.ssc_sed_flux(c["nu"], *blob, c["B"], c["R_b"], *ne, gamma=c["gamma"].copy()),
                orc.ssc_sed_flux(c["nu"], *blob, c["B"], c["R_b"], *ne, integrator="trapz_loglog",
                                 gamma=c["gamma"].copy()),
                orc.ec_blr_sed_flux(c["nu"], *blob, c["mu_s"], c["R_b"], c["L_disk"], c["xi_line"],
                                    c["epsilon_line"], c["R_line"], c["r"], *ne, gamma=c["gamma_ec"].copy(),
                                    mu=c["mu"], phi=c["phi"]),
                orc.ec_blr_sed_flux(c["nu"], *blob, c["mu_s"], c["R_b"], c["L_disk"], c["xi_line"],
                                    c["epsilon_line"], c["R_line"], c["r"], *ne, integrator="trapz_loglog",
                                    gamma=c["gamma_ec"].copy(), mu=c["mu"], phi=c["phi"]),
                orc.ec_dt_sed_flux(c["nu"], *blob, c["mu_s"], c["R_b"], c["L_disk"], c["xi_dt"],
                                   c["epsilon_dt"], c["R_dt"], c["r"], *ne, gamma=c["gamma_ec"].copy(),
                                   phi=c["phi"]),
                orc.ec_ss_disk_sed_flux(c["nu"], *blob, c["mu_s"], c["R_b"], c["M_BH"], c["L_disk"], c["eta"],
                                        c["R_in"], c["R_out"], c["r"], *ne, gamma=c["gamma_ec"].copy(),
                                        mu_size=16, phi=c["phi"]),
                orc.tau_blr(c["nu_tau"], c["z"], 1.0, c["L_disk"], c["xi_line"],
                            c["epsilon_line"], c["R_line"], c["r"], l_size=20, mu=c["mu"], phi=c["phi"])
                if seed % 2 == 0 else
                orc.tau_blr_mu_s(c["nu_tau"], c["z"], c["mu_s"], c["L_disk"], c["xi_line"], c["epsilon_line"],
                                 c["R_line"], c["r"], u_size=20, mu=c["mu"], phi=c["phi"]),
                orc.tau_dt(c["nu_tau"], c["z"], 1.0, c["L_disk"], c["xi_dt"], c["epsilon_dt"], c["R_dt"], c["r"],
                           l_size=20, phi=c["phi"]),
            ]
            continue
        Q = A.Q
        S, SSC, EC, Ab = A.agnpy.Synchrotron, A.agnpy.SynchrotronSelfCompton, A.agnpy.ExternalCompton, A.agnpy.Absorption
        bq = (c["z"], Q(c["d_L"], "cm"), c["delta_D"])
        ne = _ne_q(A, c)
        nu, nut = Q(c["nu"], "Hz"), Q(c["nu_tau"], "Hz")
        ecb = (nu, *bq, c["mu_s"], Q(c["R_b"], "cm"))
        blr = (Q(c["L_disk"], "erg s-1"), c["xi_line"], c["epsilon_line"], Q(c["R_line"], "cm"), Q(c["r"], "cm"))
        dt = (Q(c["L_disk"], "erg s-1"), c["xi_dt"], c["epsilon_dt"], Q(c["R_dt"], "cm"), Q(c["r"], "cm"))
        out += [
            _val(S.evaluate_sed_flux(nu, *bq, Q(c["B"], "G"), Q(c["R_b"], "cm"), *ne, gamma=c["gamma"].copy())),
            _val(SSC.evaluate_sed_flux(nu, *bq, Q(c["B"], "G"), Q(c["R_b"], "cm"), *ne, gamma=c["gamma"].copy())),
            _val(SSC.evaluate_sed_flux(nu, *bq, Q(c["B"], "G"), Q(c["R_b"], "cm"), *ne,
                                       integrator=A.integ("trapz_loglog"), gamma=c["gamma"].copy())),
            _val(EC.evaluate_sed_flux_blr(*ecb, *blr, *ne, gamma=c["gamma_ec"].copy(), mu=c["mu"], phi=c["phi"])),
            _val(EC.evaluate_sed_flux_blr(*ecb, *blr, *ne, integrator=A.integ("trapz_loglog"),
                                          gamma=c["gamma_ec"].copy(), mu=c["mu"], phi=c["phi"])),
            _val(EC.evaluate_sed_flux_dt(*ecb, *dt, *ne, gamma=c["gamma_ec"].copy(), phi=c["phi"])),
            _val(EC.evaluate_sed_flux_ss_disk(*ecb, Q(c["M_BH"], "g"), Q(c["L_disk"], "erg s-1"), c["eta"],
                                              Q(c["R_in"], "cm"), Q(c["R_out"], "cm"), Q(c["r"], "cm"), *ne,
                                              gamma=c["gamma_ec"].copy(), mu_size=16, phi=c["phi"])),
            _val(Ab.evaluate_tau_blr(nut, c["z"], 1.0, *blr, l_size=20, mu=c["mu"], phi=c["phi"]))
            if seed % 2 == 0 else
            _val(Ab.evaluate_tau_blr_mu_s(nut, c["z"], c["mu_s"], *blr, u_size=20, mu=c["mu"], phi=c["phi"])),
            _val(Ab.evaluate_tau_dt(nut, c["z"], 1.0, *dt, l_size=20, phi=c["phi"])),
        ]
    return np.concatenate([np.ravel(x) for x in out])


# ------------------------------------------------------------------------------------------
# config 3: gamma-gamma opacities (docs/snippets/absorption_snippet.py:13-29) and the rest of tau()
# ------------------------------------------------------------------------------------------
def _tau(side, A, idx, which, mu_s=1.0, r=None, n_nu=200):
    t = cases.case_tau(n_nu=n_nu)
    nu = t["nu"] if idx is None else t["nu"][idx]
    r = t["r"] if r is None else r
    z = t["z"]
    Ab = None if side == "orc" else A.agnpy.Absorption
    if which == "ss_disk":
        a = (cases.M_BH, cases.L_DISK, cases.ETA, 6 * cases.R_G, 200 * cases.R_G, r)
        if side == "orc":
            return orc.chunked_over_nu(orc.tau_ss_disk, nu, 10, z, mu_s, *a)
        return _chunked(lambda nn: Ab.evaluate_tau_ss_disk(
            A.Q(nn, "Hz"), z, mu_s, A.Q(a[0], "g"), A.Q(a[1], "erg s-1"), a[2], A.Q(a[3], "cm"),
            A.Q(a[4], "cm"), A.Q(a[5], "cm")), nu, 10)
    if which in ("blr", "blr_mu_s"):
        a = (cases.L_DISK, cases.XI_LINE, cases.EPS_LINE, cases.R_LINE, r)
        if side == "orc":
            fn = orc.tau_blr if which == "blr" else orc.tau_blr_mu_s
            return orc.chunked_over_nu(fn, nu, 10, z, mu_s, *a)
        fn = Ab.evaluate_tau_blr if which == "blr" else Ab.evaluate_tau_blr_mu_s
        return _chunked(lambda nn: fn(A.Q(nn, "Hz"), z, mu_s, A.Q(a[0], "erg s-1"), a[1], a[2],
                                      A.Q(a[3], "cm"), A.Q(a[4], "cm")), nu, 10)
    if which in ("dt", "dt_mu_s"):
        a = (cases.L_DISK, cases.XI_DT, cases.EPS_DT, cases.R_DT, r)
        if side == "orc":
            fn = orc.tau_dt if which == "dt" else orc.tau_dt_mu_s
            return fn(nu, z, mu_s, *a)
        fn = Ab.evaluate_tau_dt if which == "dt" else Ab.evaluate_tau_dt_mu_s
        return _val(fn(A.Q(nu, "Hz"), z, mu_s, A.Q(a[0], "erg s-1"), a[1], a[2], A.Q(a[3], "cm"),
                       A.Q(a[4], "cm")))
    if which in ("ps", "ps_mu_s"):
        if side == "orc":
            fn = orc.tau_ps_behind_blob if which == "ps" else orc.tau_ps_behind_blob_mu_s
            return fn(nu, z, mu_s, 1e-5, 2e46, r)
        fn = Ab.evaluate_tau_ps_behind_blob if which == "ps" else Ab.evaluate_tau_ps_behind_blob_mu_s
        return _val(fn(A.Q(nu, "Hz"), z, mu_s, 1e-5, A.Q(2e46, "erg s-1"), A.Q(r, "cm")))
    raise KeyError(which)


case("c3_tau_ss_disk", sub=10)(lambda side, A, idx: _tau(side, A, idx, "ss_disk"))
case("c3_tau_blr", sub=10)(lambda side, A, idx: _tau(side, A, idx, "blr"))
case("c3_tau_dt")(lambda side, A, idx: _tau(side, A, idx, "dt"))
case("tau_ss_disk_off_axis", sub=4)(
    lambda side, A, idx: _tau(side, A, idx, "ss_disk", mu_s=0.9, r=1e17, n_nu=40))
case("tau_blr_inside_shell_nudge", sub=4)(        # l grid hits R_line: targets_absorption.py:335-338
    lambda side, A, idx: _tau(side, A, idx, "blr", r=1e15, n_nu=40))
case("tau_blr_mu_s", sub=4)(lambda side, A, idx: _tau(side, A, idx, "blr_mu_s", mu_s=0.9, n_nu=40))
case("tau_blr_mu_s_outside", sub=4)(
    lambda side, A, idx: _tau(side, A, idx, "blr_mu_s", mu_s=0.5, r=2e17, n_nu=40))
case("tau_dt_mu_s")(lambda side, A, idx: _tau(side, A, idx, "dt_mu_s", mu_s=0.9, n_nu=60))
case("tau_ps_behind_blob")(lambda side, A, idx: _tau(side, A, idx, "ps", mu_s=0.9, n_nu=60))
case("tau_ps_behind_blob_mu_s")(lambda side, A, idx: _tau(side, A, idx, "ps_mu_s", mu_s=0.9, n_nu=60))


def _ref_blob(A, c, gamma_size=200, Gamma=10):
    """a reference Blob on the config-1 parameters; d_L comes from the (stub) cosmology"""
    ne = A.ne(c["ne_kind"])(A.Q(c["ne_args"][0], "cm-3"), *c["ne_args"][1:])
    return A.agnpy.Blob(R_b=A.Q(c["R_b"], "cm"), z=c["z"], delta_D=c["delta_D"],
                                         Gamma=Gamma, B=A.Q(c["B"], "G"), n_e=ne,
                                         gamma_e_size=gamma_size)


@case("tau_on_synchrotron")
def _tau_synch(side, A, idx, store=None):
    c = cases.case_c1()
    c["ne_args"] = (1e-2,) + tuple(c["ne_args"][1:])          # dense enough to absorb
    t = cases.case_tau(n_nu=40)
    if side == "orc":
        d_L = float(store["tau_on_synchrotron__d_L"])
        return orc.tau_on_synchrotron(t["nu"], c["z"], d_L, c["delta_D"], c["B"], c["R_b"],
                                      c["ne_kind"], c["ne_args"], np.logspace(1, 6, 200))
    blob = _ref_blob(A, c)
    store["tau_on_synchrotron__d_L"] = np.asarray(blob.d_L.to_value("cm"))
    ab = A.agnpy.Absorption(blob)
    return _val(ab.tau_on_synchrotron(blob, A.Q(t["nu"], "Hz")))


@case("instance_conveniences", rtol=RTOL_EXP_TAIL)
def _instances(side, A, idx, store=None):
    """the object-level API on top of the static methods: Blob -> Synchrotron / SSC /
    ExternalCompton(.set_mu, .set_phi) .sed_flux / .sed_luminosity / .sed_peak_*; Absorption(target, r, z)
    (.set_mu, .set_phi, .set_l) .tau / .absorption.  One flat vector of everything."""
    c = cases.case_c1()
    nu = np.logspace(10, 27, 18)
    nu_t = cases.case_tau(n_nu=12)["nu"]
    r_ec, r_tau, z_tau = 2e17, 1.1e16, 0.859
    n_g, n_mu, n_phi, n_l = 120, 40, 20, 30
    Gamma = 10
    mu_s = (1 - 1 / (Gamma * c["delta_D"])) / np.sqrt(1 - 1 / Gamma**2)     # blob.py:110-113
    if side == "orc":
        d_L = float(store["instance_conveniences__d_L"])
        g_blob = np.logspace(np.log10(c["ne_args"][-2]), np.log10(c["ne_args"][-1]), n_g)
        g_ext = np.logspace(1, 9, n_g)                                       # blob.py:148-151
        a = (c["z"], d_L, c["delta_D"], c["B"], c["R_b"], c["ne_kind"], c["ne_args"])
        syn = orc.synch_sed_flux(nu, *a, gamma=g_blob.copy())
        ssc = orc.ssc_sed_flux(nu, *a, gamma=g_blob.copy())
        mu, phi = np.linspace(-1, 1, n_mu), np.linspace(0, 2 * np.pi, n_phi)
        ec = orc.ec_blr_sed_flux(nu, c["z"], d_L, c["delta_D"], mu_s, c["R_b"], cases.L_DISK, cases.XI_LINE,
                                 cases.EPS_LINE, cases.R_LINE, r_ec, c["ne_kind"], c["ne_args"],
                                 gamma=g_ext.copy(), mu=mu, phi=phi)
        ec_dt = orc.ec_dt_sed_flux(nu, c["z"], d_L, c["delta_D"], mu_s, c["R_b"], cases.L_DISK, cases.XI_DT,
                                   cases.EPS_DT, cases.R_DT, r_ec, c["ne_kind"], c["ne_args"],
                                   gamma=g_ext.copy(), phi=phi)
        tau = orc.tau_blr(nu_t, z_tau, 1, cases.L_DISK, cases.XI_LINE, cases.EPS_LINE, cases.R_LINE, r_tau,
                          l_size=n_l, mu=mu, phi=phi)
        k = int(np.argmax(syn))
        return np.concatenate([syn, 4 * np.pi * d_L**2 * syn, [nu[k], syn[k]], ssc, ec, 4 * np.pi * d_L**2 * ec,
                               ec_dt, tau, np.exp(-tau)])
    blob = _ref_blob(A, c, gamma_size=n_g, Gamma=Gamma)
    store["instance_conveniences__d_L"] = np.asarray(blob.d_L.to_value("cm"))
    nuq, nutq = A.Q(nu, "Hz"), A.Q(nu_t, "Hz")
    synch = A.agnpy.Synchrotron(blob)
    ssc = A.agnpy.SynchrotronSelfCompton(blob)
    blr = A.agnpy.SphericalShellBLR(A.Q(cases.L_DISK, "erg s-1"), cases.XI_LINE, "Lyalpha", A.Q(cases.R_LINE, "cm"))
    dt = A.agnpy.RingDustTorus(A.Q(cases.L_DISK, "erg s-1"), cases.XI_DT, A.Q(cases.T_DT, "K"),
                               A.Q(cases.R_DT, "cm"))
    ec = A.agnpy.ExternalCompton(blob, blr, r=A.Q(r_ec, "cm"))
    ec.set_mu(n_mu)
    ec.set_phi(n_phi)
    ec_dt = A.agnpy.ExternalCompton(blob, dt, r=A.Q(r_ec, "cm"))
    ec_dt.set_phi(n_phi)
    ab = A.agnpy.Absorption(blr, r=A.Q(r_tau, "cm"), z=z_tau)
    ab.set_mu(n_mu)
    ab.set_phi(n_phi)
    ab.set_l(n_l)
    peak = [_val(synch.sed_peak_nu(nuq).to("Hz")), _val(synch.sed_peak_flux(nuq).to("erg cm-2 s-1"))]
    return np.concatenate([_val(synch.sed_flux(nuq)), _val(synch.sed_luminosity(nuq).to("erg s-1")),
                           np.ravel(peak), _val(ssc.sed_flux(nuq)), _val(ec.sed_flux(nuq)),
                           _val(ec.sed_luminosity(nuq).to("erg s-1")), _val(ec_dt.sed_flux(nuq)),
                           _val(ab.tau(nutq)), _val(ab.absorption(nutq))])


@case("ssc_energy_loss_rate")
def _loss(side, A, idx, store=None):
    c = cases.case_c1()
    g_eval = np.logspace(1, 7, 30)
    if side == "orc":
        d_L = float(store["ssc_energy_loss_rate__d_L"])
        return orc.ssc_electron_energy_loss_rate(g_eval, c["z"], d_L, c["delta_D"], c["B"], c["R_b"],
                                                 c["ne_kind"], c["ne_args"])
    blob = _ref_blob(A, c)
    store["ssc_energy_loss_rate__d_L"] = np.asarray(blob.d_L.to_value("cm"))
    ssc = A.agnpy.SynchrotronSelfCompton(blob)
    return _val(ssc.electron_energy_loss_rate(g_eval).to("erg s-1"))


# ------------------------------------------------------------------------------------------
# targets: integral energy densities u(r[, blob]) and thermal SEDs
# ------------------------------------------------------------------------------------------
R_EVAL = np.logspace(16, 21, 12)


def _targets(A):
    T = A.agnpy
    return dict(
        ps=T.PointSourceBehindJet(A.Q(2e46, "erg s-1"), 1e-5),
        disk=T.SSDisk(A.Q(cases.M_BH, "g"), A.Q(cases.L_DISK, "erg s-1"), cases.ETA, 6, 200,
                      R_g_units=True),
        blr=T.SphericalShellBLR(A.Q(cases.L_DISK, "erg s-1"), cases.XI_LINE, "Lyalpha",
                                A.Q(cases.R_LINE, "cm")),
        dt=T.RingDustTorus(A.Q(cases.L_DISK, "erg s-1"), cases.XI_DT, A.Q(cases.T_DT, "K")),
    )


def _u(side, A, idx, which, Gamma):
    if side == "orc":
        if which == "ps":
            return orc.u_ps_behind_jet(2e46, R_EVAL, Gamma)
        if which == "disk":
            return np.array([orc.u_ss_disk(cases.M_BH, cases.L_DISK, cases.ETA, 6 * cases.R_G,
                                           200 * cases.R_G, r, Gamma)[0] for r in R_EVAL])
        if which == "blr":
            return orc.u_blr(cases.L_DISK, cases.XI_LINE, cases.R_LINE, R_EVAL, Gamma)
        return orc.u_dt(cases.L_DISK, cases.XI_DT, cases.R_DT, R_EVAL, Gamma)
    blob = None
    if Gamma:
        blob = _ref_blob(A, cases.case_c1(), Gamma=Gamma)
    t = _targets(A)[which]
    if which == "disk":     # SSDisk.u takes one distance at a time (targets.py:441-444)
        return np.array([float(_val(t.u(A.Q(r, "cm"), blob))) for r in R_EVAL])
    return _val(t.u(A.Q(R_EVAL, "cm"), blob))


for _w in ("ps", "disk", "blr", "dt"):
    case(f"u_{_w}_stationary")(lambda side, A, idx, w=_w: _u(side, A, idx, w, None))
    case(f"u_{_w}_comoving")(lambda side, A, idx, w=_w: _u(side, A, idx, w, 20.0))

NU_BB = np.logspace(11, 17, 50)


@case("disk_multi_T_bb_sed", rtol=RTOL_EXP_TAIL)
def _bb1(side, A, idx):
    m_dot = cases.L_DISK / (cases.ETA * orc.C_LIGHT**2)
    a = (cases.Z_EC, cases.M_BH, m_dot, 6 * cases.R_G, 200 * cases.R_G, cases.D_L_EC)
    if side == "orc":
        return orc.disk_multi_T_bb_sed(NU_BB, *a, mu_s=cases.MU_S_EC)
    return _val(A.agnpy.SSDisk.evaluate_multi_T_bb_sed(
        A.Q(NU_BB, "Hz"), a[0], A.Q(a[1], "g"), A.Q(a[2], "g s-1"), A.Q(a[3], "cm"), A.Q(a[4], "cm"),
        A.Q(a[5], "cm"), mu_s=cases.MU_S_EC))


@case("disk_multi_T_bb_norm_sed", rtol=RTOL_EXP_TAIL)
def _bb2(side, A, idx):
    m_dot = cases.L_DISK / (cases.ETA * orc.C_LIGHT**2)
    a = (cases.Z_EC, cases.L_DISK, cases.M_BH, m_dot, 6 * cases.R_G, 200 * cases.R_G, cases.D_L_EC)
    if side == "orc":
        return orc.disk_multi_T_bb_norm_sed(NU_BB, *a, mu_s=cases.MU_S_EC)
    return _val(A.agnpy.SSDisk.evaluate_multi_T_bb_norm_sed(
        A.Q(NU_BB, "Hz"), a[0], A.Q(a[1], "erg s-1"), A.Q(a[2], "g"), A.Q(a[3], "g s-1"),
        A.Q(a[4], "cm"), A.Q(a[5], "cm"), A.Q(a[6], "cm"), mu_s=cases.MU_S_EC))


@case("dt_bb_sed", rtol=RTOL_EXP_TAIL)
def _bb3(side, A, idx):
    a = (cases.Z_EC, cases.T_DT, cases.R_DT, cases.D_L_EC)
    if side == "orc":
        return orc.dt_bb_sed(NU_BB, *a)
    return _val(A.agnpy.RingDustTorus.evaluate_bb_sed(
        A.Q(NU_BB, "Hz"), a[0], A.Q(a[1], "K"), A.Q(a[2], "cm"), A.Q(a[3], "cm")))


@case("dt_bb_norm_sed", rtol=RTOL_EXP_TAIL)
def _bb4(side, A, idx):
    a = (cases.Z_EC, cases.XI_DT * cases.L_DISK, cases.T_DT, cases.R_DT, cases.D_L_EC)
    if side == "orc":
        return orc.dt_bb_norm_sed(NU_BB, *a)
    return _val(A.agnpy.RingDustTorus.evaluate_bb_norm_sed(
        A.Q(NU_BB, "Hz"), a[0], A.Q(a[1], "erg s-1"), A.Q(a[2], "K"), A.Q(a[3], "cm"),
        A.Q(a[4], "cm")))


@case("target_scalars")
def _scalars(side, A, idx):
    """SSDisk / BLR / DT derived numbers: R_g, eps_line, eps_dt, R_dt (targets.py:244-252, 493-495,
    570-580) and the to_R_g_units conversion"""
    if side == "orc":
        return np.array([cases.R_G, orc.line_epsilon(), orc.dt_epsilon(cases.T_DT),
                         orc.dt_sublimation_radius(cases.L_DISK, cases.T_DT),
                         orc.to_R_g_units(1e17, cases.M_BH)])
    t = _targets(A)
    return np.array([float(t["disk"].R_g.to_value("cm")), float(t["blr"].epsilon_line),
                     float(t["dt"].epsilon_dt), float(t["dt"].R_dt.to_value("cm")),
                     float(A.to_R_g_units(1e17, cases.M_BH) if getattr(A, "is_product", False) else
                           __import__("importlib").import_module("agnpy.utils.conversion").to_R_g_units(
                               A.Q(1e17, "cm"), A.Q(cases.M_BH, "g")))])


@case("ebl_finke_2010", rtol=1e-12)
def _ebl(side, A, idx):
    """absorption/ebl_absorption.py:44-73 on the reference's own table file; NaN outside the table"""
    nu = np.logspace(24, 28, 30)
    z = 0.5
    if side == "orc":
        tab = orc.read_ebl_fits(cases.GOLDEN / "reference_data" / "ebl_models" / "frd_abs.fits.gz")
        return np.nan_to_num(orc.ebl_absorption(nu, z, *tab), nan=-1.0)
    kw = {"data_dir": cases.GOLDEN / "reference_data" / "ebl_models"} if getattr(A, "is_product", False) else {}
    return np.nan_to_num(np.asarray(A.agnpy.EBL("finke_2010", **kw).absorption(A.Q(nu, "Hz"), z)), nan=-1.0)


# ------------------------------------------------------------------------------------------
# fit-wrapper level: the reference's own scenario functions
# agnpy/fit/sherpa_wrapper.py:51-325 and gammapy_wrapper.py:161-181, 284-357, imported through
# import-level stand-ins for sherpa / gammapy (oracle/astropy_stub/{sherpa,gammapy}); outputs [n_nu, n_vec]
# ------------------------------------------------------------------------------------------
N_FIT_VECTORS = 6
EV_ERG = 1.602176634e-12


def fit_oracle(scenario, row, nu, d_L, ssa=False):
    """the wrapper's arithmetic on the oracle: parameter unpacking + log10 scaling
    (sherpa_wrapper.py:29-48), R_b = c t_var delta_D / (1+z) (:64), component sum (:66-75, 122-150,
    196-225, 276-325)"""
    cols = cases.C5_COLUMNS[scenario]
    p = dict(zip([("log10_k", "p1", "p2", "log10_gamma_b", "log10_gamma_min", "log10_gamma_max", "z",
                   "delta_D", "log10_B", "t_var", "mu_s", "log10_r", "log10_L_disk", "M_BH", "m_dot",
                   "R_in", "R_out", "xi_line", "lambda_line", "R_line", "xi_dt", "T_dt", "R_dt")[c]
                  for c in cols], row))
    ne = (10 ** p["log10_k"], p["p1"], p["p2"], 10 ** p["log10_gamma_b"], 10 ** p["log10_gamma_min"],
          10 ** p["log10_gamma_max"])
    z, delta_D, B = p["z"], p["delta_D"], 10 ** p["log10_B"]
    R_b = (orc.C_LIGHT * p["t_var"] * delta_D) / (1 + z)
    a = (nu, z, d_L, delta_D, B, R_b, "BrokenPowerLaw", ne)
    sed = orc.synch_sed_flux(*a, ssa=ssa) + orc.ssc_sed_flux(*a, ssa=ssa)
    if scenario == "ssc":
        return sed
    r, L_disk = 10 ** p["log10_r"], 10 ** p["log10_L_disk"]
    gamma_ec = np.logspace(1, 9, 300)
    sed = sed + orc.disk_multi_T_bb_norm_sed(nu, z, L_disk, p["M_BH"], p["m_dot"], p["R_in"],
                                             p["R_out"], d_L)
    if "dt" in scenario:
        eps_dt = 2.7 * (orc.K_B * p["T_dt"] / orc.MEC2)
        sed = sed + orc.dt_bb_norm_sed(nu, z, p["xi_dt"] * L_disk, p["T_dt"], p["R_dt"], d_L)
        sed = sed + orc.ec_dt_sed_flux(nu, z, d_L, delta_D, p["mu_s"], R_b, L_disk, p["xi_dt"], eps_dt,
                                       p["R_dt"], r, "BrokenPowerLaw", ne, gamma=gamma_ec)
    if "blr" in scenario:
        eps_line = orc.H_PLANCK * orc.C_LIGHT / (p["lambda_line"] * 1e-8) / orc.MEC2
        sed = sed + orc.chunked_over_nu(
            lambda nn: orc.ec_blr_sed_flux(nn, z, d_L, delta_D, p["mu_s"], R_b, L_disk, p["xi_line"],
                                           eps_line, p["R_line"], r, "BrokenPowerLaw", ne,
                                           gamma=gamma_ec), nu, 5)
    return sed


def _fit_sherpa(side, A, idx, scenario, store=None):
    pars = cases.c5_parameter_vectors(N_FIT_VECTORS)[:, cases.C5_COLUMNS[scenario]]
    nu = cases.C5_NU if idx is None else cases.C5_NU[idx]
    key = f"fit_sherpa_{scenario}__d_L"
    if side == "orc":
        d_L = store[key]
        return np.stack([fit_oracle(scenario, pars[i], nu, float(d_L[i]))
                         for i in range(N_FIT_VECTORS)], axis=1)
    import importlib
    if getattr(A, "is_product", False):
        return A.fit_sherpa(scenario, pars, nu)
    sw = importlib.import_module("agnpy.fit.sherpa_wrapper")
    fn = getattr(sw, f"_evaluate_sed_{scenario}_scenario")
    n_e = A.agnpy.BrokenPowerLaw()
    x = nu * orc.H_PLANCK / EV_ERG        # sherpa hands energies in eV
    store[key] = np.array([float(importlib.import_module(
        "astropy.coordinates").Distance(z=row[6]).to("cm").value) for row in pars])
    return np.stack([_val(fn(x.copy(), [float(v) for v in row], n_e, False)) for row in pars], axis=1)


for _sc in ("ssc", "ec_blr", "ec_dt", "ec_blr_dt"):
    case(f"fit_sherpa_{_sc}", sub=(4 if _sc == "ssc" else 12), rtol=RTOL_EXP_TAIL)(
        lambda side, A, idx, store=None, sc=_sc: _fit_sherpa(side, A, idx, sc, store))

GAMMAPY_NAMES = ("log10_k", "p1", "p2", "log10_gamma_b", "log10_gamma_min", "log10_gamma_max", "z",
                 "delta_D", "log10_B", "t_var", "mu_s", "log10_r", "log10_L_disk", "M_BH", "m_dot",
                 "R_in", "R_out", "xi_line", "lambda_line", "R_line", "xi_dt", "T_dt", "R_dt")
GAMMAPY_UNITS = {"t_var": "s", "M_BH": "g", "m_dot": "g s-1", "R_in": "cm", "R_out": "cm",
                 "lambda_line": "Angstrom", "R_line": "cm", "T_dt": "K", "R_dt": "cm"}


@case("fit_gammapy_ec_blr_dt", sub=12, rtol=RTOL_EXP_TAIL)
def _fit_gammapy(side, A, idx, store=None):
    """ExternalComptonSpectralModel(n_e, ["blr", "dt"]).evaluate(energy, **kwargs): dN/dE in
    1 / (cm2 eV s) (gammapy_wrapper.py:284-357)"""
    pars = cases.c5_parameter_vectors(3, seed=7)
    nu = cases.C5_NU if idx is None else cases.C5_NU[idx]
    energy_eV = nu * orc.H_PLANCK / EV_ERG
    if side == "orc":
        d_L = store["fit_gammapy_ec_blr_dt__d_L"]
        sed = np.stack([fit_oracle("ec_blr_dt", pars[i], nu, float(d_L[i])) for i in range(3)], axis=1)
        return sed / (energy_eV[:, None] ** 2 * EV_ERG)
    import importlib
    if getattr(A, "is_product", False):
        return A.fit_gammapy(["blr", "dt"], pars, energy_eV)
    gw = importlib.import_module("agnpy.fit.gammapy_wrapper")
    model = gw.ExternalComptonSpectralModel(A.agnpy.BrokenPowerLaw(), ["blr", "dt"])
    out = []
    for row in pars:
        kwargs = {n: A.Q(v, GAMMAPY_UNITS.get(n, "")) for n, v in zip(GAMMAPY_NAMES, row)}
        out.append(_val(model.evaluate(A.Q(energy_eV, "eV"), **kwargs)))
    store["fit_gammapy_ec_blr_dt__d_L"] = np.array([float(importlib.import_module(
        "astropy.coordinates").Distance(z=row[6]).to("cm").value) for row in pars])
    return np.stack(out, axis=1)
