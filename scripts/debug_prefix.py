import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import bench
from agnpy_b200.batch import SedBatch
g = bench.grids()
models = bench.synthetic_models(12)
a = SedBatch("ec_blr", g["nu"], "BrokenPowerLaw", gamma=g["gamma"], mu=g["mu"], phi=g["phi"], integrator="trapz_prefix")(models)
b = SedBatch("ec_blr", g["nu"], "BrokenPowerLaw", gamma=g["gamma"], mu=g["mu"], phi=g["phi"])(models)
for i in range(12):
    nz = b[i] != 0
    d = np.zeros_like(b[i]); d[nz] = np.abs(a[i][nz] / b[i][nz] - 1)
    bad = np.argsort(d)[-3:]
    print(i, "max dev", d.max(), "at nu idx", bad, d[bad], "sed/peak", b[i][bad] / b[i].max())
    if d.max() > 1e-11:
        k = int(bad[-1])
        o = bench.oracle_sed(models[i], g, nu=g["nu"][k:k + 1])[0]
        print("   model", {n: (models[i][n].tolist() if hasattr(models[i][n], 'tolist') else models[i][n]) for n in ("delta_D", "mu_s")}, "tgt", models[i]["tgt"][:5], "ne", models[i]["ne"][:6])
        print("   prefix", a[i][k], "direct", b[i][k], "oracle", o, "prefix/oracle-1", a[i][k] / o - 1, "direct/oracle-1", b[i][k] / o - 1)
