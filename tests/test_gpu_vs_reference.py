"""The CUDA path against the reference's OWN outputs, through the drop-in API (gpu).

tests/golden/ref_evaluate.npz was produced by the unmodified reference (agnpy v0.5.0) running the case
code of tests/ref_cases.py.  Here the *same case code* runs with `agnpy_b200` substituted for `agnpy`:
same class and method names, same argument order, Quantities in and out — the parity tests read like the
reference's own calls — and every frequency of every case must agree with the reference at 1e-9
(BASELINE.json north_star), exact zeros as exact zeros.

Conditioning notes (properties of the reference's formulas, see tests/ref_cases.py): results that pass
through the literal SSA attenuation are compared under the strict-reference switch at
RTOL_SSA_ATTENUATION (the reference formula itself is only that reproducible under last-bit changes of
tau), and, separately, in the default (series) mode against a cancellation-free attenuation.
"""
import numpy as np
import pytest

from agnpy_b200 import _lib
from tests import cases, ref_cases
from tests.test_gpu_parity import assert_parity

pytestmark = pytest.mark.gpu

STORE = dict(np.load(cases.GOLDEN / "ref_evaluate.npz"))
STORE.pop("__meta__", None)
RTOL = 1e-9


def _rtol(name):
    return max(RTOL, ref_cases.RTOL_SSA_ATTENUATION) if "_ssa" in name else RTOL


@pytest.mark.parametrize("name", list(ref_cases.CASES))
def test_cuda_equals_reference(name):
    A = ref_cases.ProductSide()
    with _lib.strict_reference_mode("_ssa" in name):
        got = np.asarray(ref_cases.run(name, "ref", A, dict(STORE)), dtype=np.float64)
    want = STORE[name]
    assert got.shape == want.shape
    assert_parity(got, want, rtol=_rtol(name), what=f"{name} vs reference")


@pytest.mark.parametrize("name", ["c1_synch_trapz_ssa", "c1_ssc_trapz_ssa", "shape_synch_PowerLaw_ssa"])
def test_default_ssa_mode_is_the_well_conditioned_one(name):
    """default mode: same tau (pinned above), attenuation through the series; it must agree with the
    reference wherever the reference's own formula is well conditioned (tau > 0.5 or tau < 1e-3) and
    stay within the reference's noise (3e-7) in between"""
    A = ref_cases.ProductSide()
    got = np.asarray(ref_cases.run(name, "ref", A, dict(STORE)), dtype=np.float64)
    want = STORE[name]
    assert_parity(got, want, rtol=1e-6, what=f"{name} default mode")


def test_strict_switch_changes_only_the_two_documented_functions():
    prev = _lib.strict_reference()
    try:
        assert _lib.strict_reference(True) == prev
        assert _lib.strict_reference() is True
        A = ref_cases.ProductSide()
        ec_strict = ref_cases.run("c3_ec_dt", "ref", A, dict(STORE))
        tau_strict = ref_cases.run("c3_tau_dt", "ref", A, dict(STORE))
        _lib.strict_reference(False)
        ec_default = ref_cases.run("c3_ec_dt", "ref", A, dict(STORE))
        tau_default = ref_cases.run("c3_tau_dt", "ref", A, dict(STORE))
    finally:
        _lib.strict_reference(prev)
    np.testing.assert_array_equal(ec_strict, ec_default)           # EC has no switched function
    assert_parity(tau_strict, STORE["c3_tau_dt"], rtol=RTOL, what="tau_dt strict")
    assert_parity(tau_default, STORE["c3_tau_dt"], rtol=RTOL, what="tau_dt default")


def test_drop_in_call_with_ndarray_subclass_quantities():
    """the argument objects of an astropy user -- ndarray-subclass Quantities, here the ones of the
    stand-in astropy the reference runs on, in non-CGS units -- give the same numbers as the package's
    own Quantity and as plain CGS doubles"""
    from oracle import reference_loader
    if not reference_loader.available():
        pytest.skip("needs the stand-in astropy loaded with the reference")
    reference_loader.load()
    import astropy.units as au
    import agnpy_b200 as ab
    from agnpy_b200 import u
    nu = np.logspace(15, 26, 40)
    n_e = ab.BrokenPowerLaw(1e-8 * u.Unit("cm-3"), 2.0, 3.5, 1e4, 20, 5e7)
    pars_own = (1e-8 * u.Unit("cm-3"), 2.0, 3.5, 1e4, 20, 5e7)
    pars_sub = (1e-2 * au.Unit("m-3"), 2.0, 3.5, 1e4, 20, 5e7)
    pc = 3.0856775814913673e18
    own = ab.ExternalCompton.evaluate_sed_flux_dt(
        nu * u.Hz, 1.0, 2.0e28 * u.cm, 40, 0.9997, 1e16 * u.cm, 2e46 * u.Unit("erg s-1"), 0.1, 2e-7,
        1.5e18 * u.cm, 1e18 * u.cm, n_e, *pars_own)
    sub = ab.ExternalCompton.evaluate_sed_flux_dt(
        (nu / 1e9) * au.GHz, 1.0, (2.0e28 / pc / 1e6) * au.Mpc, 40, 0.9997, 1e14 * au.m,
        2e39 * au.W, 0.1, 2e-7, (1.5e18 / pc) * au.pc, 1e16 * au.m, n_e, *pars_sub)
    plain = ab.ExternalCompton.evaluate_sed_flux_dt(
        nu, 1.0, 2.0e28, 40, 0.9997, 1e16, 2e46, 0.1, 2e-7, 1.5e18, 1e18, n_e,
        1e-8, 2.0, 3.5, 1e4, 20, 5e7)
    np.testing.assert_array_equal(own.value, np.asarray(plain.value))
    assert_parity(np.asarray(sub.value), np.asarray(own.value), rtol=1e-13, what="non-CGS ndarray-subclass inputs")
    assert np.max(own.value) > 0
