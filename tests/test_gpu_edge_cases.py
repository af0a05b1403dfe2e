"""Edge cases of the drop-in boundary on the GPU: minimal and ragged sizes, empty electron
distributions, unsorted / repeated frequencies, scalar-like inputs, random parameter sweeps
against the oracle, argument errors."""
import numpy as np
import pytest

import agnpy_b200 as ab
from agnpy_b200 import _lib, spectra, u
from oracle import agnpy_oracle as orc
from tests import cases
from tests.test_gpu_parity import assert_parity, ec_args, ne_obj, ne_tail, q_args

pytestmark = pytest.mark.gpu
ERG = u.Unit("erg s-1")


def blr_call(c, **kw):
    return ab.ExternalCompton.evaluate_sed_flux_blr(
        *ec_args(c), c["L_disk"] * ERG, c["xi_line"], c["epsilon_line"], c["R_line"] * u.cm,
        c["r"] * u.cm, *ne_tail(c), gamma=c["gamma"], mu=c["mu"], phi=c["phi"], **kw).value


def test_minimal_grids():
    """2-point grids on every axis and a single frequency"""
    c = cases.case_c2(n_nu=1, n_gamma=2, n_mu=2, n_phi=2, r=2e17)
    c["nu"] = np.array([1e22])
    c["gamma"] = np.array([1e3, 1e5])
    for integ, integrator in (("trapz", np.trapezoid), ("trapz_loglog", ab.trapz_loglog),
                              ("trapz", ab.trapz_prefix)):
        got = blr_call(c, integrator=integrator)
        ref = cases.ORACLE_EC["blr"](c, integ)
        assert got.shape == (1,)
        if ref[0] == 0:
            assert got[0] == 0
        else:
            assert abs(got[0] / ref[0] - 1) < 1e-9


def test_unsorted_and_repeated_frequencies_keep_shape():
    c = cases.case_c2(n_nu=16, r=1e18)
    rng = np.random.default_rng(3)
    nu = rng.permutation(np.concatenate([c["nu"], c["nu"][:4]]))
    c2 = dict(c, nu=nu)
    got = blr_call(c2)
    assert_parity(got, cases.ORACLE_EC["blr"](c2), what="unsorted nu")
    nu2d = nu.reshape(4, 5)
    got2 = ab.ExternalCompton.evaluate_sed_flux_blr(
        nu2d * u.Hz, c["z"], c["d_L"] * u.cm, c["delta_D"], c["mu_s"], c["R_b"] * u.cm,
        c["L_disk"] * ERG, c["xi_line"], c["epsilon_line"], c["R_line"] * u.cm, c["r"] * u.cm,
        *ne_tail(c), gamma=c["gamma"], mu=c["mu"], phi=c["phi"])
    assert got2.shape == (4, 5)
    np.testing.assert_array_equal(got2.value.ravel(), got)


def test_electrons_outside_the_gamma_grid_give_exact_zeros():
    """n_e = 0 on the whole integration grid: every entry point returns exact zeros"""
    args = (1e-8, 2.0, 1e10, 1e11)            # above logspace(1, 9)
    ne = spectra.PowerLaw(*args)
    nu = np.logspace(10, 28, 12) * u.Hz
    common = (nu, 0.1, 1e27 * u.cm, 10.0)
    for integrator in (np.trapezoid, ab.trapz_loglog):
        s = ab.Synchrotron.evaluate_sed_flux(*common, 1 * u.G, 1e16 * u.cm, ne, *q_args(args),
                                             integrator=integrator, ssa=True)
        assert np.all(s.value == 0)
        s = ab.SynchrotronSelfCompton.evaluate_sed_flux(*common, 1 * u.G, 1e16 * u.cm, ne,
                                                        *q_args(args), integrator=integrator)
        assert np.all(s.value == 0)
        s = ab.ExternalCompton.evaluate_sed_flux_dt(*common, 0.99, 1e16 * u.cm, 1e45 * ERG, 0.1, 1e-6,
                                                    1e18 * u.cm, 1e17 * u.cm, ne, *q_args(args),
                                                    integrator=integrator)
        assert np.all(s.value == 0)
    s = ab.ExternalCompton.evaluate_sed_flux_blr(*common, 0.99, 1e16 * u.cm, 1e45 * ERG, 0.1, 2e-5,
                                                 1e17 * u.cm, 1e18 * u.cm, ne, *q_args(args),
                                                 integrator=ab.trapz_prefix)
    assert np.all(s.value == 0)


def test_frequencies_beyond_the_kinematic_limit_are_exact_zeros():
    c = cases.case_c2(n_nu=12, r=1e18)
    c["nu"] = np.logspace(31, 34, 12)          # eps_s > gamma_max delta_D: nothing can scatter there
    for integrator in (np.trapezoid, ab.trapz_loglog, ab.trapz_prefix):
        assert np.all(blr_call(c, integrator=integrator) == 0)
    t = ab.Absorption.evaluate_tau_blr(np.logspace(18, 21, 8) * u.Hz, 0.5, 1.0, 2e46 * ERG, 0.024,
                                       cases.EPS_LINE, 1e17 * u.cm, 1e16 * u.cm)
    assert np.all(t == 0)                       # below the pair-production threshold


@pytest.mark.parametrize("seed", range(6))
def test_random_parameter_sweep_against_oracle(seed):
    """random physical parameters in (and beyond) the fit ranges of fit/core.py:92-157"""
    rng = np.random.default_rng(1000 + seed)
    kind = ["PowerLaw", "BrokenPowerLaw", "LogParabola", "ExpCutoffPowerLaw",
            "ExpCutoffBrokenPowerLaw"][seed % 5]
    gmin, gmax = 10 ** rng.uniform(0.5, 2.5), 10 ** rng.uniform(5, 8)
    shape = {"PowerLaw": (rng.uniform(1.5, 3.5),),
             "BrokenPowerLaw": (rng.uniform(1.5, 2.5), rng.uniform(2.6, 4.5), 10 ** rng.uniform(3, 5)),
             "LogParabola": (rng.uniform(1.5, 3), rng.uniform(0.05, 0.6), 10 ** rng.uniform(2.5, 4.5)),
             "ExpCutoffPowerLaw": (rng.uniform(1.5, 3), 10 ** rng.uniform(3.5, 6)),
             "ExpCutoffBrokenPowerLaw": (rng.uniform(1.5, 2.5), rng.uniform(2.6, 4), 10 ** rng.uniform(4, 6),
                                         10 ** rng.uniform(3, 4.5))}[kind]
    args = (10 ** rng.uniform(-9, -5),) + shape + (gmin, gmax)
    z, delta = rng.uniform(0.01, 2.5), rng.uniform(2, 50)
    Gamma = rng.uniform(max(2.0, delta / 2), delta * 1.2)
    beta = np.sqrt(1 - 1 / Gamma**2)
    mu_s = float(np.clip((1 - 1 / (Gamma * delta)) / beta, -1, 1))
    d_L, R_b, B = 10 ** rng.uniform(26.5, 28.5), 10 ** rng.uniform(15, 17), 10 ** rng.uniform(-2, 1)
    nu = np.logspace(rng.uniform(8, 12), rng.uniform(24, 30), 28)
    gamma = np.logspace(np.log10(gmin), np.log10(gmax), int(rng.integers(40, 260)))
    ne = ne_obj(kind, args)
    integ, integrator = (("trapz", np.trapezoid), ("trapz_loglog", ab.trapz_loglog))[seed % 2]
    got = ab.Synchrotron.evaluate_sed_flux(nu * u.Hz, z, d_L * u.cm, delta, B * u.G, R_b * u.cm, ne,
                                           *q_args(args), integrator=integrator, gamma=gamma)
    assert_parity(got.value, orc.synch_sed_flux(nu, z, d_L, delta, B, R_b, kind, args,
                                                integrator=integ, gamma=gamma), what="sweep synch")
    got = ab.SynchrotronSelfCompton.evaluate_sed_flux(nu * u.Hz, z, d_L * u.cm, delta, B * u.G,
                                                      R_b * u.cm, ne, *q_args(args),
                                                      integrator=integrator, gamma=gamma)
    assert_parity(got.value, orc.ssc_sed_flux(nu, z, d_L, delta, B, R_b, kind, args,
                                              integrator=integ, gamma=gamma), what="sweep ssc")
    L, xi, R_line, r = 10 ** rng.uniform(44, 47), rng.uniform(0.01, 0.3), 10 ** rng.uniform(16.5, 18), 10 ** rng.uniform(15.5, 19.5)
    g_ec = np.logspace(1, 9, int(rng.integers(60, 220)))
    mu, phi = np.linspace(-1, 1, int(rng.integers(11, 60))), np.linspace(0, 2 * np.pi, int(rng.integers(7, 40)))
    got = ab.ExternalCompton.evaluate_sed_flux_blr(
        nu * u.Hz, z, d_L * u.cm, delta, mu_s, R_b * u.cm, L * ERG, xi, cases.EPS_LINE, R_line * u.cm,
        r * u.cm, ne, *q_args(args), integrator=integrator, gamma=g_ec, mu=mu, phi=phi)
    ref = orc.ec_blr_sed_flux(nu, z, d_L, delta, mu_s, R_b, L, xi, cases.EPS_LINE, R_line, r, kind, args,
                              integrator=integ, gamma=g_ec, mu=mu, phi=phi)
    assert_parity(got.value, ref, what="sweep ec blr")
    if integ == "trapz":
        got = ab.ExternalCompton.evaluate_sed_flux_blr(
            nu * u.Hz, z, d_L * u.cm, delta, mu_s, R_b * u.cm, L * ERG, xi, cases.EPS_LINE,
            R_line * u.cm, r * u.cm, ne, *q_args(args), integrator=ab.trapz_prefix, gamma=g_ec, mu=mu, phi=phi)
        assert_parity(got.value, ref, what="sweep ec blr prefix")
    T_dt = rng.uniform(300, 2000)
    eps_dt, R_dt = orc.dt_epsilon(T_dt), orc.dt_sublimation_radius(L, T_dt)
    got = ab.ExternalCompton.evaluate_sed_flux_dt(
        nu * u.Hz, z, d_L * u.cm, delta, mu_s, R_b * u.cm, L * ERG, 0.1, eps_dt, R_dt * u.cm, r * u.cm,
        ne, *q_args(args), integrator=integrator, gamma=g_ec, phi=phi)
    assert_parity(got.value, orc.ec_dt_sed_flux(nu, z, d_L, delta, mu_s, R_b, L, 0.1, eps_dt, R_dt, r, kind,
                                                args, integrator=integ, gamma=g_ec, phi=phi), what="sweep ec dt")
    nu_tau = np.logspace(24, 30, 20)
    got = ab.Absorption.evaluate_tau_blr(nu_tau * u.Hz, z, 1.0, L * ERG, xi, cases.EPS_LINE, R_line * u.cm,
                                         r * u.cm, l_size=33, mu=mu, phi=phi)
    assert_parity(got, orc.tau_blr(nu_tau, z, 1.0, L, xi, cases.EPS_LINE, R_line, r, l_size=33, mu=mu, phi=phi),
                  what="sweep tau blr")
    got = ab.Absorption.evaluate_tau_dt(nu_tau * u.Hz, z, 1.0, L * ERG, 0.1, eps_dt, R_dt * u.cm, r * u.cm,
                                        l_size=41, phi=phi)
    assert_parity(got, orc.tau_dt(nu_tau, z, 1.0, L, 0.1, eps_dt, R_dt, r, l_size=41, phi=phi), what="sweep tau dt")


def test_argument_errors():
    ne = spectra.PowerLaw(1e-8, 2.0, 10, 1e5)
    nu = np.logspace(10, 20, 5) * u.Hz
    with pytest.raises(ValueError):           # unknown integrator: no host fallback
        ab.Synchrotron.evaluate_sed_flux(nu, 0.1, 1e27 * u.cm, 10, 1 * u.G, 1e16 * u.cm, ne,
                                         *ne.parameters, integrator=np.sum)
    with pytest.raises(TypeError):            # wrong physical dimension
        ab.Synchrotron.evaluate_sed_flux(nu, 0.1, 1e27 * u.Hz, 10, 1 * u.G, 1e16 * u.cm, ne, *ne.parameters)
    with pytest.raises(ValueError):           # wrong number of spectral parameters
        ab.Synchrotron.evaluate_sed_flux(nu, 0.1, 1e27 * u.cm, 10, 1 * u.G, 1e16 * u.cm, ne, 1e-8, 2.0)
    with pytest.raises(ValueError):           # descending gamma grid
        ab.Synchrotron.evaluate_sed_flux(nu, 0.1, 1e27 * u.cm, 10, 1 * u.G, 1e16 * u.cm, ne,
                                         *ne.parameters, gamma=np.logspace(5, 1, 50))
    # C ABI argument checks
    L = _lib.lib()
    assert L.agnpy_b200_ec_sed_host(99, 1, None, 0, None, 1, None, 2, None, 2, None, 2, None, 0, None) != 0
    assert b"ec_sed" in L.agnpy_b200_last_error()
    models = _lib.make_models(1)
    out = np.zeros(3)
    rc = L.agnpy_b200_tau_host(_lib.TAU_BLR, 1, models.ctypes.data, out.ctypes.data, 3, None, 0,
                               out.ctypes.data, 3, out.ctypes.data, 3, out.ctypes.data)
    assert rc == 1 and b"mu grid" in L.agnpy_b200_last_error()


def test_units_are_converted_not_assumed():
    """parsec / Mpc / W / kg inputs give the same SED as their CGS equivalents"""
    c = cases.case_c3_dt(n_nu=20)
    ne = ne_tail(c)
    a = ab.ExternalCompton.evaluate_sed_flux_dt(
        c["nu"] * u.Hz, c["z"], c["d_L"] * u.cm, c["delta_D"], c["mu_s"], c["R_b"] * u.cm,
        c["L_disk"] * ERG, c["xi_dt"], c["epsilon_dt"], c["R_dt"] * u.cm, c["r"] * u.cm, *ne,
        gamma=c["gamma"], phi=c["phi"])
    pc = 3.0856775814913673e18
    b = ab.ExternalCompton.evaluate_sed_flux_dt(
        c["nu"] / 1e9 * u.GHz, c["z"], c["d_L"] / (1e6 * pc) * u.Mpc, c["delta_D"], c["mu_s"],
        c["R_b"] / 100 * u.m, c["L_disk"] / 1e7 * u.W, c["xi_dt"], c["epsilon_dt"], c["R_dt"] / pc * u.pc,
        c["r"] / pc * u.pc, ne[0], ne[1].to("m-3"), *ne[2:], gamma=c["gamma"], phi=c["phi"])
    np.testing.assert_allclose(a.value, b.value, rtol=1e-12)


def test_model_axis_chunking(monkeypatch):
    """the model axis is processed in scratch-limited chunks: force tiny chunks and check that
    every process gives the same rows as one unchunked call"""
    import bench
    from agnpy_b200.batch import SedBatch
    from agnpy_b200 import _core
    g = bench.grids()
    nu = g["nu"][::10]
    models = bench.synthetic_models(11)
    kw = dict(gamma=g["gamma"][::2], mu=g["mu"][::5], phi=g["phi"][::5])
    tmodels = _lib.make_models(11)
    for i in range(11):
        _core.model_row(tmodels, i, z=0.3, mu_s=1.0, tgt=[cases.L_DISK, 0.024, cases.EPS_LINE, 1e17, 1e16 * (1 + i)])
    ref = {}
    for proc in ("ec_blr", "ssc", "synch", "tau_blr"):
        m = tmodels if proc == "tau_blr" else models
        ref[proc] = SedBatch(proc, nu if proc != "tau_blr" else cases.case_tau(12)["nu"], "BrokenPowerLaw", **kw)(m)
    monkeypatch.setenv("AGNPY_B200_CHUNK_BYTES", "300000")      # 1-3 models per chunk
    for proc in ("ec_blr", "ssc", "synch", "tau_blr"):
        m = tmodels if proc == "tau_blr" else models
        got = SedBatch(proc, nu if proc != "tau_blr" else cases.case_tau(12)["nu"], "BrokenPowerLaw", **kw)(m)
        np.testing.assert_array_equal(got, ref[proc])
        assert np.count_nonzero(got) > 0


@pytest.mark.parametrize("grid", ["fine", "linear", "coarse"])
def test_unusual_gamma_grids(grid):
    """very fine grids put many gamma points inside the mask guard bands (exact fall-back path);
    linear and very coarse grids change the search geometry"""
    gamma = {"fine": np.logspace(2, 5, 1500), "linear": np.linspace(30.0, 3e5, 400),
             "coarse": np.logspace(1, 9, 12)}[grid]
    c = cases.case_c2(n_nu=14, r=2e17, n_mu=21, n_phi=11)
    c["gamma"] = gamma
    for integ, integrator in (("trapz", np.trapezoid), ("trapz_loglog", ab.trapz_loglog), ("trapz", ab.trapz_prefix)):
        assert_parity(blr_call(c, integrator=integrator), cases.ORACLE_EC["blr"](c, integ), what=f"ec {grid} {integ}")
    c1 = cases.case_c1(n_nu=16)
    g1 = {"fine": np.logspace(1, 4, 1200), "linear": np.linspace(10.0, 1e5, 300), "coarse": np.logspace(1, 6, 9)}[grid]
    for integ, integrator in (("trapz", np.trapezoid), ("trapz_loglog", ab.trapz_loglog)):
        got = ab.SynchrotronSelfCompton.evaluate_sed_flux(
            c1["nu"] * u.Hz, c1["z"], c1["d_L"] * u.cm, c1["delta_D"], c1["B"] * u.G, c1["R_b"] * u.cm,
            ne_obj(c1["ne_kind"], c1["ne_args"]), *q_args(c1["ne_args"]), integrator=integrator, gamma=g1)
        ref = orc.ssc_sed_flux(c1["nu"], c1["z"], c1["d_L"], c1["delta_D"], c1["B"], c1["R_b"], c1["ne_kind"],
                               c1["ne_args"], integrator=integ, gamma=g1)
        assert_parity(got.value, ref, what=f"ssc {grid} {integ}")


# ---- the np.trapz polynomial EC path: frequency chunks, sorted pair segments, plan switches ------
@pytest.mark.parametrize("l_nu", ["1", "3", "8"])
@pytest.mark.parametrize("order", ["ascending", "descending", "shuffled"])
def test_ec_poly_frequency_chunks(monkeypatch, l_nu, order):
    """a block walks L_nu frequencies and advances its index window from one to the next:
    any chunk length and any frequency order must give the single-frequency result"""
    c = cases.case_c2(n_nu=29, r=3e17, n_gamma=150, n_mu=40, n_phi=24)   # 960 pairs, ragged chunks
    nu = c["nu"]
    if order == "descending":
        nu = nu[::-1].copy()
    elif order == "shuffled":
        nu = np.random.default_rng(11).permutation(nu)
    c = dict(c, nu=nu)
    monkeypatch.setenv("AGNPY_B200_EC_LNU", l_nu)
    got = blr_call(c)
    assert_parity(got, cases.ORACLE_EC["blr"](c), what=f"L_nu={l_nu} {order}")
    monkeypatch.setenv("AGNPY_B200_EC_LNU", "1")
    np.testing.assert_array_equal(blr_call(c), got)     # chunking never changes a bit


@pytest.mark.parametrize("n_mu,n_phi", [(16, 32), (17, 32), (31, 17), (70, 31), (130, 67)])
def test_ec_poly_pair_counts(monkeypatch, n_mu, n_phi):
    """pair counts around the plan switches: 512 (v4 kernel below, polynomial path from there on),
    one partly filled 1024-pair sort segment, several segments, tiles of 4 and of 5 pairs per thread"""
    monkeypatch.setenv("AGNPY_B200_EC_LNU", "4")
    c = cases.case_c2(n_nu=9, r=2e17, n_gamma=90, n_mu=n_mu, n_phi=n_phi)
    assert_parity(blr_call(c), cases.ORACLE_EC["blr"](c), what=f"{n_mu}x{n_phi} pairs")


def test_ec_poly_off_axis_strong_phi_dependence(monkeypatch):
    """mu_s far from 1: c varies strongly with phi, the sorted order interleaves the mu rows"""
    monkeypatch.setenv("AGNPY_B200_EC_LNU", "8")
    c = cases.case_c2(n_nu=17, r=1e18, n_gamma=120, n_mu=48, n_phi=40)
    c = dict(c, mu_s=0.3)
    assert_parity(blr_call(c), cases.ORACLE_EC["blr"](c), what="mu_s = 0.3")


def test_ec_poly_long_gamma_grid(monkeypatch):
    """1000 gamma points: 32 KB rows, one frequency per block whatever the knob asks for"""
    monkeypatch.setenv("AGNPY_B200_EC_LNU", "8")
    c = cases.case_c2(n_nu=5, r=5e17, n_gamma=1000, n_mu=24, n_phi=24)
    assert_parity(blr_call(c), cases.ORACLE_EC["blr"](c), what="n_gamma = 1000")


@pytest.mark.parametrize("plan", ["128,5", "128,4", "64,8", "128,8"])
def test_ec_poly_tile_shapes_agree(monkeypatch, plan):
    """every instantiated tile shape of the polynomial kernel (threads x pairs per thread; the default is
    chosen by padding, the others stay reachable through AGNPY_B200_EC_PLAN for tuning) gives the oracle's
    result, and the shapes agree with each other to rounding (the pair order inside a thread differs)"""
    c = cases.case_c2(n_nu=23, r=2e17, n_gamma=130, n_mu=60, n_phi=44)       # 1320 folded pairs
    monkeypatch.delenv("AGNPY_B200_EC_PLAN", raising=False)
    default = blr_call(c)
    monkeypatch.setenv("AGNPY_B200_EC_PLAN", plan)
    got = blr_call(c)
    assert_parity(got, cases.ORACLE_EC["blr"](c), what=f"plan {plan}")
    assert_parity(got, default, rtol=1e-13, what=f"plan {plan} vs default")


@pytest.mark.parametrize("integrator", ["trapz", "trapz_prefix", "trapz_loglog"])
def test_repeated_launches_are_bit_identical(integrator):
    """the kernels synchronise through mbarriers, shuffles, ballots and shared-memory hand-overs and
    are written to be deterministic (fixed reduction orders, no atomics on the data path): 25 launches
    of the same batch -- enough to fill the GPU several times over, so that blocks are scheduled
    differently from launch to launch -- must agree to the last bit (a shared-memory race or a missing
    barrier shows up as run-to-run noise; compute-sanitizer racecheck is not available on the pool)"""
    import torch
    from agnpy_b200.batch import SedBatch
    import bench
    g = bench.grids()
    n = 8 if integrator == "trapz_loglog" else 48
    models = bench.synthetic_models(n)
    plan = SedBatch("ec_blr", g["nu"], "BrokenPowerLaw", gamma=g["gamma"], mu=g["mu"], phi=g["phi"],
                    integrator=integrator)
    first = plan(models)
    assert np.all(np.isfinite(first)) and first.max() > 0
    for _ in range(25 if integrator != "trapz_loglog" else 6):
        np.testing.assert_array_equal(plan(models), first)
    c1 = cases.case_c1()
    ssc_models = _lib.make_models(64)
    rng = np.random.default_rng(3)
    from agnpy_b200 import _core
    for i in range(64):
        _core.model_row(ssc_models, i, z=c1["z"], d_L=c1["d_L"], delta_D=rng.uniform(5, 30), mu_s=1.0,
                        B=10 ** rng.uniform(-1.5, 0.5), R_b=10 ** rng.uniform(15.5, 16.5),
                        ne=[10 ** rng.uniform(-9, -7), rng.uniform(1.5, 2.5), rng.uniform(2.6, 4),
                            10 ** rng.uniform(3.5, 5), 10.0, 1e6, 0, 0])
    ssc = SedBatch("ssc", c1["nu"], "BrokenPowerLaw", gamma=c1["gamma"],
                   integrator="trapz" if integrator == "trapz_prefix" else integrator)
    first = ssc(ssc_models)
    for _ in range(25):
        np.testing.assert_array_equal(ssc(ssc_models), first)


# ---- the np.trapz SSC path: gamma slots, seed grids that are not log-uniform ----------------------
@pytest.mark.parametrize("n_gamma", [2, 31, 33, 300])
def test_ssc_gamma_slot_counts(n_gamma):
    c = cases.case_c1(n_nu=23)
    gamma = np.logspace(1, 6, n_gamma)
    got = ab.SynchrotronSelfCompton.evaluate_sed_flux(
        c["nu"] * u.Hz, c["z"], c["d_L"] * u.cm, c["delta_D"], c["B"] * u.G, c["R_b"] * u.cm,
        ne_obj(c["ne_kind"], c["ne_args"]), *q_args(c["ne_args"]), gamma=gamma).value
    ref = orc.ssc_sed_flux(c["nu"], c["z"], c["d_L"], c["delta_D"], c["B"], c["R_b"], c["ne_kind"],
                           c["ne_args"], gamma=gamma)
    assert_parity(got, ref, what=f"ssc n_gamma={n_gamma}")


def test_ssc_batch_on_irregular_seed_grid():
    """the seed-range search starts from a log-uniform guess; an irregular seed grid must fall back
    to the plain search and still match (batched entry point, custom nu_seed)"""
    from agnpy_b200 import _core
    from agnpy_b200.batch import SedBatch
    c = cases.case_c1(n_nu=19)
    rng = np.random.default_rng(5)
    nu_seed = np.sort(10 ** rng.uniform(5, 30, 137))
    models = _lib.make_models(3)
    for i in range(3):
        _core.model_row(models, i, z=c["z"], d_L=c["d_L"], delta_D=c["delta_D"] * (1 + i), B=c["B"],
                        R_b=c["R_b"], ne=list(c["ne_args"]) + [0, 0])
    got = SedBatch("ssc", c["nu"], c["ne_kind"], gamma=c["gamma"], nu_seed=nu_seed)(models)
    for i in range(3):
        ref = orc.ssc_sed_flux(c["nu"], c["z"], c["d_L"], c["delta_D"] * (1 + i), c["B"], c["R_b"],
                               c["ne_kind"], c["ne_args"], gamma=c["gamma"], nu_seed=nu_seed)
        assert_parity(got[i], ref, what=f"irregular seed grid, model {i}")


# ---- folded phi axis (AGNPY_INTEG_FOLD_PHI) -------------------------------------------------------
def test_phi_fold_flag_host_logic_and_agreement(monkeypatch):
    """mirror-symmetric phi grids are folded by default; the unfolded evaluation (knob) agrees to
    rounding and both match the oracle; odd and even point counts, a grid symmetric about 0"""
    from agnpy_b200 import grids
    assert grids.phi_fold_flag(np.linspace(0, 2 * np.pi, 50)) == _lib.INTEG_FOLD_PHI
    assert grids.phi_fold_flag(np.linspace(-np.pi, np.pi, 33)) == _lib.INTEG_FOLD_PHI
    assert grids.phi_fold_flag(np.linspace(0.1, 2 * np.pi, 50)) == 0
    assert grids.phi_fold_flag(np.sort(np.random.default_rng(0).uniform(0, 6, 30))) == 0
    for n_phi, lo, hi in ((50, 0.0, 2 * np.pi), (51, 0.0, 2 * np.pi), (32, -np.pi, np.pi)):
        c = cases.case_c2(n_nu=13, r=4e17, n_gamma=110, n_mu=40, n_phi=n_phi)
        c = dict(c, phi=np.linspace(lo, hi, n_phi))
        monkeypatch.delenv("AGNPY_B200_EC_NOFOLD", raising=False)
        folded = blr_call(c)
        monkeypatch.setenv("AGNPY_B200_EC_NOFOLD", "1")
        plain = blr_call(c)
        monkeypatch.delenv("AGNPY_B200_EC_NOFOLD", raising=False)
        ref = cases.ORACLE_EC["blr"](c)
        assert_parity(folded, ref, what=f"folded n_phi={n_phi}")
        assert_parity(plain, ref, what=f"unfolded n_phi={n_phi}")
        nz = ref != 0
        assert np.max(np.abs(folded[nz] / plain[nz] - 1)) < 1e-12


def test_phi_fold_keeps_the_mask_decision_of_each_mirror_partner():
    """The one thing the last bit of cos(phi) can change is on which side of a table entry G_k a
    pair's threshold c falls.  Build that case on purpose: one mirror partner is moved by 1e-7 rad
    (its c by ~1e-9 relative) and a frequency is placed so that G_k(nu) lies between the two
    thresholds -- the reference masks gamma_k for one partner and not for the other.  A folded
    evaluation that let the representative decide for both would be off by ~1e-6 of the SED."""
    from agnpy_b200 import _core, spectra
    c = cases.case_c2(n_nu=3, r=3e17, n_gamma=120, n_mu=24, n_phi=48)   # 1152 pairs, 576 folded
    phi = c["phi"].copy()
    i_phi, i_mu, k = 9, 15, 70
    phi[48 - 1 - i_phi] += 1e-7
    mu_star = orc.mu_star_shell(c["mu"][i_mu], c["R_line"], c["r"])
    cs = [2 * c["epsilon_line"] * (1 - orc.cos_psi(c["mu_s"], mu_star, p)) for p in (phi[i_phi], phi[48 - 1 - i_phi])]
    assert 1e-10 < abs(cs[1] / cs[0] - 1) < 1e-7
    G = np.sqrt(cs[0] * cs[1])                       # between the two thresholds
    g = c["gamma"][k]
    eps_s = G * g * g / (1 + G * g)                  # G = eps_s / (gamma (gamma - eps_s))
    nu_star = eps_s * orc.MEC2 / orc.H_PLANCK / (1 + c["z"])
    c = dict(c, phi=phi, nu=np.array([nu_star / 3, nu_star, nu_star * 3]))
    # the reference really decides differently for the two partners at (nu_star, gamma_k)
    gmin = [orc.gamma_min_compton(orc.nu_to_epsilon_prime(nu_star, c["z"]), c["epsilon_line"], c["mu_s"],
                                  mu_star, p) for p in (phi[i_phi], phi[48 - 1 - i_phi])]
    assert (g >= gmin[0]) != (g >= gmin[1])
    ref = cases.ORACLE_EC["blr"](c)
    kind, par = spectra.describe(ne_obj(c["ne_kind"], c["ne_args"]), q_args(c["ne_args"]))
    models = _lib.make_models(1)
    _core.model_row(models, 0, z=c["z"], d_L=c["d_L"], delta_D=c["delta_D"], mu_s=c["mu_s"], R_b=c["R_b"],
                    ne=par, tgt=[c["L_disk"], c["xi_line"], c["epsilon_line"], c["R_line"], c["r"]])
    for flag in (0, _lib.INTEG_FOLD_PHI):            # the caller vouches for the (almost) symmetric grid
        got = _core.ec_sed_host(_lib.EC_BLR, models, kind, c["nu"], c["gamma"], c["mu"], c["mu"].size, phi,
                                None, _lib.INTEG_TRAPZ | flag)[0]
        assert_parity(got, ref, what=f"mirror partners straddling a table entry, flag={flag:#x}")
