"""Short SSC / tau_blr / synch runs for ncu (batched, BASELINE grids)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from agnpy_b200 import _core, _lib
from agnpy_b200.batch import SedBatch
from tests import cases

torch.cuda.set_device(0)
rng = np.random.default_rng(1)
c1 = cases.case_c1()
n = 64
models = _lib.make_models(n)
for i in range(n):
    _core.model_row(models, i, z=c1["z"], d_L=c1["d_L"], delta_D=rng.uniform(5, 30), mu_s=1.0,
                    B=10 ** rng.uniform(-1.5, 0.5), R_b=10 ** rng.uniform(15.5, 16.5),
                    ne=[10 ** rng.uniform(-9, -7), rng.uniform(1.5, 2.5), rng.uniform(2.6, 4),
                        10 ** rng.uniform(3.5, 5), 10.0, 1e6, 0, 0],
                    tgt=[cases.L_DISK, 0.024, cases.EPS_LINE, 1e17, 10 ** rng.uniform(16, 17)])
for proc, nu, kw in (("ssc", c1["nu"], dict(gamma=c1["gamma"])),
                     ("tau_blr", cases.case_tau()["nu"], {}),
                     ("synch", c1["nu"], dict(gamma=c1["gamma"], ssa=True))):
    plan = SedBatch(proc, nu, "BrokenPowerLaw", **kw)
    m = models if proc != "tau_blr" else models[:8]
    for _ in range(2):
        out = plan(m)
    print(proc, out.shape, float(np.nanmax(out)))
