"""GPU parity for SURVEY 8(f) rank 4: target energy densities u(r[, blob]) and the EBL table,
through the drop-in classes -> C ABI, against the oracle at rtol 1e-9."""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import agnpy_b200 as ab
from agnpy_b200 import targets, u
from oracle import agnpy_oracle as orc
from tests import cases

EBL_DIR = Path(__file__).parent / "golden" / "reference_data" / "ebl_models"
L_DISK = 1.512e46
ERG_S = u.Unit("erg s-1")


class _Blob:
    def __init__(self, Gamma):
        self.Gamma = Gamma
        self.Beta = np.sqrt(1 - 1 / Gamma**2)


def _close(got, ref, rtol=1e-9):
    got = np.asarray(got.value if hasattr(got, "value") else got)
    assert got.shape == np.shape(ref)
    assert np.all(np.isfinite(got))
    assert np.max(np.abs(got / ref - 1)) < rtol, np.max(np.abs(got / ref - 1))


@pytest.mark.parametrize("Gamma", [None, 10.0, 37.5])
def test_u_point_source_and_torus(Gamma):
    blob = None if Gamma is None else _Blob(Gamma)
    r = np.logspace(16, 23, 57)
    ps = targets.PointSourceBehindJet(1e46 * ERG_S, 1e-5)
    _close(ps.u(r * u.cm, blob), orc.u_ps_behind_jet(1e46, r, Gamma))
    dt = targets.RingDustTorus(L_DISK * ERG_S, 0.1, 1000 * u.K)
    R_dt = orc.dt_sublimation_radius(L_DISK, 1000.0)
    _close(dt.u(r * u.cm, blob), orc.u_dt(L_DISK, 0.1, R_dt, r, Gamma))


@pytest.mark.parametrize("Gamma", [None, 10.0, 37.5])
def test_u_blr(Gamma):
    blob = None if Gamma is None else _Blob(Gamma)
    blr = targets.SphericalShellBLR(L_DISK * ERG_S, 0.1, "Lyalpha", 1e17 * u.cm)
    r = np.logspace(16, 20, 10)                    # the reference's own test grid (crosses R_line)
    _close(blr.u(r * u.cm, blob), orc.u_blr(L_DISK, 0.1, 1e17, r, Gamma))
    r2 = np.logspace(15, 23, 2 * 3 * 7).reshape(2, 3, 7)   # shape is kept
    got = blr.u(r2 * u.cm, blob)
    _close(got, orc.u_blr(L_DISK, 0.1, 1e17, r2.ravel(), Gamma).reshape(r2.shape))


@pytest.mark.parametrize("Gamma", [None, 10.0])
def test_u_ss_disk(Gamma):
    blob = None if Gamma is None else _Blob(Gamma)
    M_BH = 1.2e9 * orc.M_SUN
    R_g = orc.G_NEWTON * M_BH / orc.C_LIGHT**2
    disk = targets.SSDisk(M_BH * u.g, L_DISK * ERG_S, 1 / 12, 6, 200, R_g_units=True)
    r = np.logspace(15.5, 20, 25)
    _close(disk.u(r * u.cm, blob), orc.u_ss_disk(M_BH, L_DISK, 1 / 12, 6 * R_g, 200 * R_g, r, Gamma), rtol=1e-9)


def test_u_cmb_host():
    cmb = targets.CMB(z=1)
    assert np.isclose(cmb.u().value, 6.67945605e-12, rtol=1e-5)
    assert np.isclose(cmb.u(_Blob(10.0)).value, orc.u_cmb(cmb.u_0.value, 10.0), rtol=1e-14)


def test_u_batched_c_abi():
    """device flavour semantics through the host entry point: many models x many r, per-model Gamma"""
    from agnpy_b200 import _core, _lib
    rng = np.random.default_rng(3)
    n = 37
    models = _lib.make_models(n)
    Ls, Rs, Gs = 10 ** rng.uniform(45, 47, n), 10 ** rng.uniform(16.5, 17.5, n), rng.uniform(2, 40, n)
    for i in range(n):
        _core.model_row(models, i, tgt=[Ls[i], 0.1, cases.EPS_LINE, Rs[i]])
    r = np.logspace(15, 22, 129)
    got = _core.target_u_host(_lib.EC_BLR, models, r, Gs, 50)
    for i in range(n):
        ref = orc.u_blr(Ls[i], 0.1, Rs[i], r, Gs[i])
        assert np.max(np.abs(got[i] / ref - 1)) < 1e-9


def test_ebl_absorption_and_apply():
    ebl = ab.EBL("finke_2010", data_dir=EBL_DIR)
    tab = orc.read_ebl_fits(EBL_DIR / "frd_abs.fits.gz")
    nu = np.logspace(23.5, 28.3, 211)
    for z in (0.0, 0.03, 0.5, 1.0, 2.345, 4.9):
        got = ebl.absorption(nu * u.Hz, z)
        ref = orc.ebl_absorption(nu, z, *tab)
        assert got.shape == nu.shape
        assert np.array_equal(np.isnan(got), np.isnan(ref))
        ok = ~np.isnan(ref)
        assert ok.sum() > 100
        assert np.allclose(got[ok], ref[ok], rtol=1e-12, atol=1e-300)
    # outside the table in redshift / energy: NaN like the reference
    assert np.all(np.isnan(ebl.absorption(nu * u.Hz, 7.0)))
    assert np.isnan(ebl.absorption(np.array([1e10]) * u.Hz, 0.5)[0])
    # in place on a batch with one redshift per model
    zs = np.array([0.1, 0.5, 1.5])
    sed = np.abs(np.random.default_rng(0).normal(size=(3, nu.size))) + 1.0
    want = sed * np.stack([orc.ebl_absorption(nu, z, *tab) for z in zs])
    out = ebl.apply(sed, nu * u.Hz, zs)
    assert out is sed
    ok = ~np.isnan(want)
    assert np.array_equal(np.isnan(sed), ~ok) and np.allclose(sed[ok], want[ok], rtol=1e-12)


def test_fit_batch_with_ebl():
    from agnpy_b200.fit import FitSedBatch
    ebl = ab.EBL("finke_2010", data_dir=EBL_DIR)
    tab = orc.read_ebl_fits(EBL_DIR / "frd_abs.fits.gz")
    nu = np.logspace(24, 27, 40)
    pars = np.array([[-8.0, 2.0, 3.5, 4.0, np.log10(20), np.log10(5e7), 0.5, 20.0, -0.5, 86400.0],
                     [-7.5, 2.2, 3.2, 4.5, np.log10(20), np.log10(5e7), 1.0, 30.0, -1.0, 86400.0]])
    plain = FitSedBatch(nu, "BrokenPowerLaw", scenario="ssc")(pars, 1e28)
    with_ebl = FitSedBatch(nu, "BrokenPowerLaw", scenario="ssc", ebl=ebl)(pars, 1e28)
    for i, z in enumerate((0.5, 1.0)):
        a = orc.ebl_absorption(nu, z, *tab)
        assert np.allclose(with_ebl[i], plain[i] * a, rtol=1e-12, atol=0)
