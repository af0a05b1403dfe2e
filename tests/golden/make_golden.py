"""Regenerate the committed golden fixtures.  Run ONLY in the build container, where the
reference checkout exists at /root/reference (it does not exist on the GPU box; nothing
in tests/, smoke() or bench.py reads it at run time).

Two kinds of fixtures are written next to this script:

1. `ref_kernels.npz` -- outputs of the reference's OWN pure-numpy hot-path functions
   (`agnpy/utils/math.py`, `agnpy/utils/geometry.py`, `agnpy/compton/kernels.py`), loaded
   verbatim by path through a stub `astropy.units` (the full package cannot be imported:
   astropy/matplotlib are not installed), evaluated on seeded inputs.  The oracle's
   restatement of those functions must reproduce them bit-for-bit.
2. `reference_data/` -- the digitised literature / jetset curves the reference's own tests
   compare against (small CSV data vectors from `agnpy/data/reference_{seds,taus}`; data,
   not source code), so the oracle can be checked at the reference's tolerances anywhere.

Usage:  python tests/golden/make_golden.py
"""
import importlib.util
import shutil
import sys
import types
from pathlib import Path

import numpy as np

REF = Path("/root/reference/agnpy")
HERE = Path(__file__).resolve().parent


def load_reference_modules():
    """stub astropy.units (only `Hz` is touched at import time,
    agnpy/utils/math.py:7) and empty parent packages so relative imports resolve."""
    if "astropy" in sys.modules and not getattr(sys.modules["astropy"], "_stub", False):
        raise RuntimeError("a real astropy is imported; run in a fresh interpreter")
    astropy = types.ModuleType("astropy")
    astropy._stub = True
    units = types.ModuleType("astropy.units")
    units.Hz = 1.0
    astropy.units = units
    sys.modules["astropy"], sys.modules["astropy.units"] = astropy, units
    for name, path in [("agnpy", REF), ("agnpy.utils", REF / "utils"),
                       ("agnpy.compton", REF / "compton")]:
        m = types.ModuleType(name)
        m.__path__ = [str(path)]
        sys.modules[name] = m

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        return m

    ref_math = load("agnpy.utils.math", REF / "utils/math.py")
    ref_geom = load("agnpy.utils.geometry", REF / "utils/geometry.py")
    ref_kern = load("agnpy.compton.kernels", REF / "compton/kernels.py")
    return ref_math, ref_geom, ref_kern


def make_kernel_fixture():
    ref_math, ref_geom, ref_kern = load_reference_modules()
    rng = np.random.default_rng(20261019)
    out = {}
    with np.errstate(all="ignore"):
        # ---- geometry ---------------------------------------------------
        mu = np.linspace(-1, 1, 37)
        phi = np.linspace(0, 2 * np.pi, 23)
        mu_s = np.array([1.0, 0.9996873, 0.5, -0.3])
        MU_S, MU, PHI = np.meshgrid(mu_s, mu, phi, indexing="ij")
        out["geom_mu_s"], out["geom_mu"], out["geom_phi"] = mu_s, mu, phi
        out["cos_psi"] = ref_geom.cos_psi(MU_S, MU, PHI)
        R_re, r = 1e17, np.array([1e16, 0.99e17, 2e17, 1e19])
        out["shell_R"], out["shell_r"] = np.array(R_re), r
        out["x_re_shell"] = np.stack([ref_geom.x_re_shell(mu, R_re, ri) for ri in r])
        out["mu_star_shell"] = np.stack([ref_geom.mu_star_shell(mu, R_re, ri) for ri in r])
        out["x_re_ring"] = ref_geom.x_re_ring(1.5652e19, r)
        # ---- off-axis geometry (utils/geometry.py:58-183) -------------------
        phi_re = np.linspace(0, 2 * np.pi, 17)[None, :, None]
        mu_re = np.linspace(-1, 1, 13)[:, None, None]
        uu = (np.logspace(-5, 5, 21) * 3e17)[None, None, :]
        out["off_phi_re"], out["off_mu_re"], out["off_u"] = phi_re.ravel(), mu_re.ravel(), uu.ravel()
        for tag, mus in (("a", 0.9999), ("b", 0.5), ("c", 0.0)):
            out[f"x_ring_{tag}"] = ref_geom.x_re_ring_mu_s(1e18, 3e17, phi_re, uu, mus)
            ph, mu_ = ref_geom.phi_mu_re_ring(1e18, 3e17, phi_re, uu, mus)
            out[f"phi_ring_{tag}"], out[f"mu_ring_{tag}"] = ph, mu_
            out[f"x_shell_{tag}"] = ref_geom.x_re_shell_mu_s(1e17, 3e17, phi_re, mu_re, uu, mus)
            ph, mu_ = ref_geom.phi_mu_re_shell(1e17, 3e17, phi_re, mu_re, uu, mus)
            out[f"phi_shell_{tag}"], out[f"mu_shell_{tag}"] = ph, mu_
        # ---- Compton kernels -------------------------------------------
        gamma = np.logspace(1, 9, 61)
        eps = np.logspace(-12, 0, 25)
        eps_s = np.logspace(-8, 10, 41)
        g3, e3, es3 = ref_math.axes_reshaper(gamma, eps, eps_s)
        out["iso_gamma"], out["iso_eps"], out["iso_eps_s"] = gamma, eps, eps_s
        out["isotropic_kernel"] = ref_kern.isotropic_kernel(g3, e3, es3)
        mu_c = np.linspace(-1, 1, 11)
        phi_c = np.linspace(0, 2 * np.pi, 9)
        g4, m4, p4, es4 = ref_math.axes_reshaper(gamma, mu_c, phi_c, eps_s)
        out["ck_mu"], out["ck_phi"] = mu_c, phi_c
        out["ck_eps0"], out["ck_mu_s"] = np.array(1.9958625603e-05), np.array(0.9996873)
        out["compton_kernel"] = ref_kern.compton_kernel(
            g4, es4, 1.9958625603e-05, 0.9996873, m4, p4)
        out["gamma_min"] = ref_kern.get_gamma_min(es4, 1.9958625603e-05, 0.9996873, m4, p4)
        # ---- trapz_loglog ----------------------------------------------
        x = np.logspace(0.3, 6.1, 53)
        y = np.stack([rng.uniform(0.5, 2.0) * x ** rng.uniform(-4, 2) for _ in range(12)], 1)
        y[10:14, 3] = 0.0                       # zeros must kill their intervals
        y[:, 5] = 7.0 * x ** -1.0               # m == -1 branch
        out["ll_x"], out["ll_y"] = x, y
        out["trapz_loglog"] = ref_math.trapz_loglog(y, x, axis=0)
        out["np_trapz"] = (np.trapezoid if hasattr(np, "trapezoid") else np.trapz)(y, x, axis=0)
        # ---- default grids ---------------------------------------------
        out["gamma_e_to_integrate"] = ref_math.gamma_e_to_integrate
        out["nu_to_integrate"] = np.asarray(ref_math.nu_to_integrate)
        out["mu_to_integrate"] = ref_math.mu_to_integrate
        out["phi_to_integrate"] = ref_math.phi_to_integrate
    np.savez_compressed(HERE / "ref_kernels.npz", **out)
    print("wrote", HERE / "ref_kernels.npz", {k: np.shape(v) for k, v in out.items()})


GOLDEN_FILES = [
    "reference_seds/dermer_menon_2009/figure_7_4/synchrotron_gamma_max_1e5.txt",
    "reference_seds/dermer_menon_2009/figure_7_4/synchrotron_gamma_max_1e7.txt",
    "reference_seds/dermer_menon_2009/figure_7_4/ssc_gamma_max_1e5.txt",
    "reference_seds/dermer_menon_2009/figure_7_4/ssc_gamma_max_1e7.txt",
    "reference_seds/finke_2016/figure_8/ec_disk_r_1e17.txt",
    "reference_seds/finke_2016/figure_8/ec_disk_r_1e18.txt",
    "reference_seds/finke_2016/figure_10/ec_blr_r_1e18.txt",
    "reference_seds/finke_2016/figure_10/ec_blr_r_1e19.txt",
    "reference_seds/finke_2016/figure_11/ec_dt_r_1e20.txt",
    "reference_seds/finke_2016/figure_11/ec_dt_r_1e21.txt",
    "reference_seds/jetset/data/synch_ssa_pwl_jetset_1.1.2.txt",
    "reference_seds/jetset/data/synch_ssa_bpwl_jetset_1.1.2.txt",
    "reference_seds/jetset/data/synch_ssa_lp_jetset_1.1.2.txt",
    "reference_seds/jetset/data/ec_cmb_bpwl_jetset_1.1.2.txt",
    "reference_taus/finke_2016/figure_14_left/tau_BLR_Ly_alpha_r_1e-1_R_Ly_alpha.txt",
    "reference_taus/finke_2016/figure_14_left/tau_BLR_Ly_alpha_r_1e0.5_R_Ly_alpha.txt",
    "reference_taus/finke_2016/figure_14_left/tau_BLR_Ly_alpha_r_1e1_R_Ly_alpha.txt",
    "reference_taus/finke_2016/figure_14_left/tau_DT_r_1e-1_R_Ly_alpha.txt",
    "reference_taus/finke_2016/figure_14_left/tau_DT_r_1e0_R_Ly_alpha.txt",
    "reference_taus/finke_2016/figure_14_left/tau_DT_r_1e1_R_Ly_alpha.txt",
    "reference_taus/finke_2016/figure_14_left/tau_DT_r_1e2_R_Ly_alpha.txt",
    "reference_taus/finke_2016/figure_14_left/tau_SSdisk_r_1e0_R_Ly_alpha.txt",
]


def copy_reference_data():
    for rel in GOLDEN_FILES:
        dst = HERE / "reference_data" / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(REF / "data" / rel, dst)
    print("copied", len(GOLDEN_FILES), "golden data files")


if __name__ == "__main__":
    make_kernel_fixture()
    copy_reference_data()
