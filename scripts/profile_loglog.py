"""Short trapz_loglog runs for ncu: EC-BLR (config 2, 16 models), SSC and synchrotron (config 1, 64 models)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bench
from agnpy_b200 import _core, _lib
from agnpy_b200.batch import SedBatch
from tests import cases

torch.cuda.set_device(0)
g = bench.grids()
plan = SedBatch("ec_blr", g["nu"], "BrokenPowerLaw", gamma=g["gamma"], mu=g["mu"], phi=g["phi"],
                integrator="trapz_loglog")
for _ in range(2):
    out = plan(bench.synthetic_models(16))
print("ec_blr loglog", out.shape, float(np.nanmax(out)))
rng = np.random.default_rng(1)
c1 = cases.case_c1()
n = 64
models = _lib.make_models(n)
for i in range(n):
    _core.model_row(models, i, z=c1["z"], d_L=c1["d_L"], delta_D=rng.uniform(5, 30), mu_s=1.0,
                    B=10 ** rng.uniform(-1.5, 0.5), R_b=10 ** rng.uniform(15.5, 16.5),
                    ne=[10 ** rng.uniform(-9, -7), rng.uniform(1.5, 2.5), rng.uniform(2.6, 4),
                        10 ** rng.uniform(3.5, 5), 10.0, 1e6, 0, 0])
for proc in ("ssc", "synch"):
    plan = SedBatch(proc, c1["nu"], "BrokenPowerLaw", gamma=c1["gamma"], integrator="trapz_loglog")
    for _ in range(2):
        out = plan(models)
    print(proc, "loglog", out.shape, float(np.nanmax(out)))
