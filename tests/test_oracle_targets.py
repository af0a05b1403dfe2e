"""Oracle pins for SURVEY 8(f) rank 4: the integral energy densities u(r[, blob]) against the
known-answer vectors of the reference's own tests (agnpy/targets/tests/test_targets.py:50-90,
223-292, 326-372; `blob_test = Blob()` has Gamma = 10), and the EBL table reader / interpolation
against the reference's table file and its interpolation test (absorption/tests/
test_absorption.py:536-590: interpolating at the table's own energies reproduces the row of the
nearest redshift)."""
from pathlib import Path

import numpy as np
import pytest

from oracle import agnpy_oracle as orc

L_DISK = 1.512 * 1e46            # test_targets.py:27
GAMMA_BLOB = 10.0                # emission_regions/blob.py: Blob() default
EBL_FILE = Path(__file__).parent / "golden" / "reference_data" / "ebl_models" / "frd_abs.fits.gz"


def test_u_cmb():
    u_0 = 7.5657 * 1e-15 * 2.72548**4 * 2**4      # targets.py:165-167 at z = 1
    assert np.isclose(orc.u_cmb(u_0), 6.67945605e-12, rtol=1e-5, atol=0)
    assert np.isclose(orc.u_cmb(u_0, GAMMA_BLOB), 8.88373221e-10, rtol=1e-3, atol=0)


def test_u_point_source():
    r = np.asarray([1e18, 1e19, 1e20])
    assert np.allclose(orc.u_ps_behind_jet(1e46, r), [2.65441873e-02, 2.65441873e-04, 2.65441873e-06],
                       rtol=1e-3, atol=0)
    assert np.allclose(orc.u_ps_behind_jet(1e46, r, GAMMA_BLOB),
                       [6.6693519e-05, 6.6693519e-07, 6.6693519e-09], rtol=1e-3, atol=0)


def test_u_blr():
    r = np.logspace(16, 20, 10)
    ref = [4.02698710e-01, 4.12267268e-01, 5.49297935e-01, 9.36951182e-02, 1.12734943e-02,
           1.44410780e-03, 1.86318218e-04, 2.40606696e-05, 3.10750072e-06, 4.01348246e-07]
    ref_co = [1.35145224e01, 1.39362806e01, 1.98574615e01, 7.22733332e-02, 2.50276530e-04,
              5.60160671e-06, 4.97415343e-07, 6.09348663e-08, 7.81583912e-09, 1.00855244e-09]
    assert np.allclose(orc.u_blr(L_DISK, 0.1, 1e17, r), ref, rtol=1e-8, atol=0)
    assert np.allclose(orc.u_blr(L_DISK, 0.1, 1e17, r, GAMMA_BLOB), ref_co, rtol=1e-8, atol=0)
    # r >> R_line: the shell tends to a point source of the same luminosity (test_targets.py:244-292)
    far = np.logspace(19, 23, 10)
    assert np.allclose(orc.u_blr(L_DISK, 0.1, 1e17, far), orc.u_ps_behind_jet(0.1 * L_DISK, far), rtol=1e-2)
    assert np.allclose(orc.u_blr(L_DISK, 0.1, 1e17, far, GAMMA_BLOB),
                       orc.u_ps_behind_jet(0.1 * L_DISK, far, GAMMA_BLOB), rtol=1e-1)


def test_u_dt():
    R_dt = orc.dt_sublimation_radius(L_DISK, 1000.0)
    assert np.isclose(R_dt, 1.361e19, rtol=1e-3)
    r = np.logspace(17, 23, 10)
    ref = [2.16675545e-05, 2.16435491e-05, 2.11389842e-05, 1.40715277e-05, 1.71541601e-06,
           8.61241559e-08, 4.01273788e-09, 1.86287690e-10, 8.64677950e-12, 4.01348104e-13]
    ref_co = [2.13519004e-03, 2.02003750e-03, 1.50733234e-03, 2.37521472e-04, 3.50603715e-07,
              4.21027697e-10, 1.04563603e-11, 4.68861573e-13, 2.17274347e-14, 1.00842240e-15]
    assert np.allclose(orc.u_dt(L_DISK, 0.1, R_dt, r), ref, rtol=1e-8, atol=0)
    assert np.allclose(orc.u_dt(L_DISK, 0.1, R_dt, r, GAMMA_BLOB), ref_co, rtol=1e-8, atol=0)


def test_u_ss_disk_far_field():
    """the reference has no known-answer vector for SSDisk.u (PARITY UNPINNED for this function);
    sanity: positive, decreasing, and ~ r^-2 (point-like) far above the disk"""
    M_BH = 1.2e9 * orc.M_SUN
    R_g = orc.G_NEWTON * M_BH / orc.C_LIGHT**2
    r = np.logspace(17.5, 20, 6)
    u = orc.u_ss_disk(M_BH, L_DISK, 1 / 12, 6 * R_g, 200 * R_g, r)
    assert np.all(u > 0) and np.all(np.diff(u) < 0)
    slope = np.diff(np.log(u[-3:])) / np.diff(np.log(r[-3:]))
    assert np.allclose(slope, -2.0, atol=0.01)
    assert np.all(orc.u_ss_disk(M_BH, L_DISK, 1 / 12, 6 * R_g, 200 * R_g, r, GAMMA_BLOB) > 0)


def test_ebl_table_and_interpolation():
    energy_ref, z_ref, values_ref = orc.read_ebl_fits(EBL_FILE)
    assert energy_ref.shape == (99,) and z_ref.shape == (500,) and values_ref.shape == (500, 99)
    assert energy_ref.dtype == np.float32 and np.all(np.diff(energy_ref) > 0) and np.all(np.diff(z_ref) > 0)
    assert np.all((values_ref >= 0) & (values_ref <= 1))
    nu_ref = energy_ref[1:-1].astype(float) * orc.KEV_ERG / orc.H_PLANCK
    for z in (0.3, 0.5, 1.0, 1.5):
        got = orc.ebl_absorption(nu_ref, z, energy_ref, z_ref, values_ref)
        ref = values_ref[np.abs(z - z_ref).argmin()][1:-1]
        assert np.allclose(got, ref, rtol=1e-4, atol=1e-7)
    # outside the table: NaN (bounds_error=False)
    assert np.all(np.isnan(orc.ebl_absorption(nu_ref[:3], 10.0, energy_ref, z_ref, values_ref)))


def test_product_ebl_reader_matches_oracle_reader():
    from agnpy_b200.ebl import read_ebl_table
    a, b = orc.read_ebl_fits(EBL_FILE), read_ebl_table(EBL_FILE)
    for x, y in zip(a, b):
        assert x.dtype == y.dtype and np.array_equal(x, y)
