"""Golden outputs of the CPU oracle at the BASELINE configs (full default grids), so the GPU
parity tests can check full-size results without paying the oracle's minutes on the GPU box.
Run in the build container:  python tests/golden/make_oracle_outputs.py
(The oracle itself is pinned against the reference by tests/test_oracle_*.py.)"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import agnpy_oracle as orc  # noqa: E402
from tests import cases  # noqa: E402

out = {}
# C1: tutorial blob, synch + SSC, both integrators
c = cases.case_c1()
for integ in ("trapz", "trapz_loglog"):
    out[f"c1_synch_{integ}"] = orc.synch_sed_flux(c["nu"], c["z"], c["d_L"], c["delta_D"], c["B"],
                                                  c["R_b"], c["ne_kind"], c["ne_args"],
                                                  integrator=integ, gamma=c["gamma"])
    out[f"c1_ssc_{integ}"] = orc.ssc_sed_flux(c["nu"], c["z"], c["d_L"], c["delta_D"], c["B"],
                                              c["R_b"], c["ne_kind"], c["ne_args"],
                                              integrator=integ, gamma=c["gamma"])
# C2: EC on the BLR, 200 nu, default grids, three distances
for r in (1e16, 2e17, 1e18):
    c2 = cases.case_c2(r=r)
    out[f"c2_ec_blr_r{r:g}"] = np.concatenate(
        [cases.ORACLE_EC["blr"](dict(c2, nu=c2["nu"][i:i + 20])) for i in range(0, 200, 20)])
# C3: EC on the DT + opacities along the path
c3 = cases.case_c3_dt()
out["c3_ec_dt"] = cases.ORACLE_EC["dt"](c3)
t = cases.case_tau()
out["c3_tau_dt"] = orc.tau_dt(t["nu"], t["z"], 1.0, 2e46, 0.1, cases.EPS_DT, cases.R_DT, t["r"])
out["c3_tau_blr"] = orc.chunked_over_nu(orc.tau_blr, t["nu"], 20, t["z"], 1.0, 2e46, 0.024,
                                        cases.EPS_LINE, 1e17, t["r"])
out["c3_tau_ss_disk"] = orc.chunked_over_nu(orc.tau_ss_disk, t["nu"], 20, t["z"], 1.0, cases.M_BH,
                                            2e46, 1 / 12, 6 * cases.R_G, 200 * cases.R_G, t["r"])
np.savez_compressed(Path(__file__).resolve().parent / "oracle_outputs.npz", **out)
print({k: (v.shape, float(np.max(v))) for k, v in out.items()})
