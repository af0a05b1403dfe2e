"""Short runs of every kernel family for ncu (round 2): EC-BLR direct + prefix (config 2, 64 models),
EC-BLR / SSC / synchrotron trapz_loglog, SSC / synchrotron np.trapz (config 1), tau_blr (config 3)."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bench
from agnpy_b200 import _core, _lib
from agnpy_b200.batch import SedBatch
from tests import cases

torch.cuda.set_device(0)
g = bench.grids()
N_MODELS = int(os.environ.get("PROFILE_MODELS", "64"))   # 256 = the batch of the bench step
ec_models = bench.synthetic_models(N_MODELS)
for integ in ("trapz", "trapz_prefix", "trapz_loglog"):
    plan = SedBatch("ec_blr", g["nu"], "BrokenPowerLaw", gamma=g["gamma"], mu=g["mu"], phi=g["phi"],
                    integrator=integ)
    out = plan(ec_models[:16] if integ == "trapz_loglog" else ec_models)
    print("ec_blr", integ, out.shape, float(np.nanmax(out)))
rng = np.random.default_rng(1)
c1 = cases.case_c1()
n = N_MODELS
models = _lib.make_models(n)
for i in range(n):
    _core.model_row(models, i, z=c1["z"], d_L=c1["d_L"], delta_D=rng.uniform(5, 30), mu_s=1.0,
                    B=10 ** rng.uniform(-1.5, 0.5), R_b=10 ** rng.uniform(15.5, 16.5),
                    ne=[10 ** rng.uniform(-9, -7), rng.uniform(1.5, 2.5), rng.uniform(2.6, 4),
                        10 ** rng.uniform(3.5, 5), 10.0, 1e6, 0, 0],
                    tgt=[cases.L_DISK, 0.024, cases.EPS_LINE, 1e17, 10 ** rng.uniform(16, 17)])
for proc, integ in (("ssc", "trapz"), ("ssc", "trapz_loglog"), ("synch", "trapz"), ("synch", "trapz_loglog")):
    plan = SedBatch(proc, c1["nu"], "BrokenPowerLaw", gamma=c1["gamma"], integrator=integ, ssa=(proc == "synch"))
    out = plan(models)
    print(proc, integ, out.shape, float(np.nanmax(out)))
plan = SedBatch("tau_blr", cases.case_tau()["nu"], "BrokenPowerLaw")
out = plan(models[:8])
print("tau_blr", out.shape, float(np.nanmax(out)))
