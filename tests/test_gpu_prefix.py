"""The opt-in suffix-sum evaluation of np.trapz over gamma (AGNPY_INTEG_TRAPZ_PREFIX) in the
external-Compton kernels: same 1e-9 bar against the oracle's np.trapz results, and agreement
with the direct kernel to re-association level."""
import numpy as np
import pytest

import agnpy_b200 as ab
from agnpy_b200 import u
from agnpy_b200.batch import SedBatch
from oracle import agnpy_oracle as orc
from tests import cases
from tests.test_gpu_parity import assert_parity, ec_args, ne_obj, ne_tail, q_args

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    return np.load(cases.GOLDEN / "oracle_outputs.npz")


@pytest.mark.parametrize("r", [1e16, 2e17, 1e18])
def test_prefix_blr_config2(gold, r):
    c = cases.case_c2(r=r)
    call = lambda integ: ab.ExternalCompton.evaluate_sed_flux_blr(
        *ec_args(c), c["L_disk"] * u.Unit("erg s-1"), c["xi_line"], c["epsilon_line"],
        c["R_line"] * u.cm, c["r"] * u.cm, *ne_tail(c), integrator=integ, gamma=c["gamma"],
        mu=c["mu"], phi=c["phi"]).value
    fast, direct = call(ab.trapz_prefix), call(ab.grids.np_trapz)
    assert_parity(fast, gold[f"c2_ec_blr_r{r:g}"], what=f"prefix C2 r={r:g}")
    assert_parity(fast, direct, rtol=1e-10, what="prefix vs direct")


def test_prefix_other_targets(gold):
    c = cases.case_c3_dt()
    got = ab.ExternalCompton.evaluate_sed_flux_dt(
        *ec_args(c), c["L_disk"] * u.Unit("erg s-1"), c["xi_dt"], c["epsilon_dt"], c["R_dt"] * u.cm,
        c["r"] * u.cm, *ne_tail(c), integrator="trapz_prefix", gamma=c["gamma"], phi=c["phi"])
    assert_parity(got.value, gold["c3_ec_dt"], what="prefix dt")
    c = cases.case_disk(n_nu=40, r=1e17)
    got = ab.ExternalCompton.evaluate_sed_flux_ss_disk(
        *ec_args(c), c["M_BH"] * u.g, c["L_disk"] * u.Unit("erg s-1"), c["eta"], c["R_in"] * u.cm,
        c["R_out"] * u.cm, c["r"] * u.cm, *ne_tail(c), integrator=ab.trapz_prefix, gamma=c["gamma"],
        mu_size=c["mu_size"], phi=c["phi"])
    assert_parity(got.value, cases.ORACLE_EC["ss_disk"](c), what="prefix disk")
    b = cases.blob_ec()
    nu = np.logspace(15, 29, 60)
    got = ab.ExternalCompton.evaluate_sed_flux_ps_behind_jet(
        nu * u.Hz, b["z"], b["d_L"] * u.cm, b["delta_D"], b["mu_s"], b["R_b"] * u.cm, cases.EPS_LINE,
        0.024 * 2e46 * u.Unit("erg s-1"), 1e19 * u.cm, ne_obj(b["ne_kind"], b["ne_args"]),
        *q_args(b["ne_args"]), integrator=ab.trapz_prefix)
    ref = orc.ec_ps_behind_jet_sed_flux(nu, b["z"], b["d_L"], b["delta_D"], b["mu_s"], b["R_b"],
                                        cases.EPS_LINE, 0.024 * 2e46, 1e19, b["ne_kind"], b["ne_args"])
    assert_parity(got.value, ref, what="prefix ps")
    c = cases.case_c2(n_nu=30, r=1e18)
    c["mu_s"] = 1.0
    got = ab.ExternalCompton.evaluate_sed_flux_blr(
        *ec_args(c), c["L_disk"] * u.Unit("erg s-1"), c["xi_line"], c["epsilon_line"],
        c["R_line"] * u.cm, c["r"] * u.cm, *ne_tail(c), integrator=ab.trapz_prefix,
        gamma=c["gamma"], mu=c["mu"], phi=c["phi"])
    assert_parity(got.value, cases.ORACLE_EC["blr"](c), what="prefix blr mu_s=1")
    c = cases.case_c2(n_nu=7, r=3e17, n_gamma=53, n_mu=17, n_phi=9)
    got = ab.ExternalCompton.evaluate_sed_flux_blr(
        *ec_args(c), c["L_disk"] * u.Unit("erg s-1"), c["xi_line"], c["epsilon_line"],
        c["R_line"] * u.cm, c["r"] * u.cm, *ne_tail(c), integrator=ab.trapz_prefix,
        gamma=c["gamma"], mu=c["mu"], phi=c["phi"])
    assert_parity(got.value, cases.ORACLE_EC["blr"](c), what="prefix ragged")


def test_prefix_batch_and_large_gamma():
    import bench
    g = bench.grids()
    models = bench.synthetic_models(12)
    a = SedBatch("ec_blr", g["nu"], "BrokenPowerLaw", gamma=g["gamma"], mu=g["mu"], phi=g["phi"],
                 integrator="trapz_prefix")(models)
    b = SedBatch("ec_blr", g["nu"], "BrokenPowerLaw", gamma=g["gamma"], mu=g["mu"], phi=g["phi"])(models)
    for i in range(12):
        assert_parity(a[i], b[i], rtol=1e-10, what=f"prefix batch {i}")
    c = cases.case_c2(n_nu=12, n_gamma=1000, n_mu=60, n_phi=30, r=2e17)   # config-4-like gamma axis
    call = lambda integ: ab.ExternalCompton.evaluate_sed_flux_blr(
        *ec_args(c), c["L_disk"] * u.Unit("erg s-1"), c["xi_line"], c["epsilon_line"],
        c["R_line"] * u.cm, c["r"] * u.cm, *ne_tail(c), integrator=integ, gamma=c["gamma"],
        mu=c["mu"], phi=c["phi"]).value
    assert_parity(call(ab.trapz_prefix), cases.ORACLE_EC["blr"](c), what="prefix 1000 gamma")
