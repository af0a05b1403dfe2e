"""Small-batch latency of the fit-level model with and without the CUDA graph (one GPU)."""
import json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from agnpy_b200.fit import FitSedBatch
from tests import cases

torch.cuda.set_device(0)
nu = np.logspace(10, 28, 100)
rows = []
for n in (1, 8, 32, 128):
    r = np.random.default_rng(n)
    p = np.zeros((n, 23)); d = r.uniform(10, 40, n)
    p[:, 0], p[:, 1], p[:, 2], p[:, 3] = r.uniform(-9, -7, n), r.uniform(1.5, 2.5, n), r.uniform(3, 4, n), r.uniform(3.5, 5, n)
    p[:, 4], p[:, 5], p[:, 6], p[:, 7], p[:, 8], p[:, 9] = np.log10(20), np.log10(5e7), 1.0, d, r.uniform(-1.5, 0, n), 86400.0
    p[:, 10] = (1 - 1 / d**2) / np.sqrt(1 - 1 / d**2)
    p[:, 11], p[:, 12] = r.uniform(17, 19, n), np.log10(2e46)
    p[:, 13:] = [cases.M_BH, 1e26, 6 * cases.R_G, 200 * cases.R_G, 0.024, 1215.67, 1e17, 0.1, 1e3, cases.R_DT]
    model = FitSedBatch(nu, "BrokenPowerLaw", scenario="ec_blr_dt")
    g = model.graphed(n)
    out = {}
    for label, fn in (("plain", model), ("graph", g)):
        for _ in range(5): fn(p, cases.D_L_EC)
        t0 = time.perf_counter()
        reps = 50
        for _ in range(reps): fn(p, cases.D_L_EC)
        out[label + "_us_per_call"] = (time.perf_counter() - t0) / reps * 1e6
    out.update(n_vectors=n, identical=bool(np.array_equal(model(p, cases.D_L_EC), g(p, cases.D_L_EC))))
    rows.append(out); print(json.dumps(out), flush=True)
