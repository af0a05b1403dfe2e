"""Time every kernel family at the BASELINE configs (device-resident, CUDA events) and print a
table: ms per call, dominant-kernel ms, SED/s, integrand points/s, nominal flops fraction."""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bench
from agnpy_b200 import _core, _lib
from agnpy_b200.batch import SedBatch
from tests import cases

torch.cuda.set_device(0)
peak = _lib.fp64_peak(200.0)
rows = []

def models_for(process, n):
    rng = np.random.default_rng(7)
    models = _lib.make_models(n)
    b = cases.blob_ec()
    for i in range(n):
        tgt = {"ec_blr": [cases.L_DISK, 0.024, cases.EPS_LINE, 1e17, 10 ** rng.uniform(17, 19)],
               "ec_dt": [cases.L_DISK, 0.1, cases.EPS_DT, cases.R_DT, 10 ** rng.uniform(18, 20)],
               "ec_ss_disk": [cases.M_BH, cases.L_DISK, 1 / 12, 6 * cases.R_G, 200 * cases.R_G, 10 ** rng.uniform(17, 18)],
               "ec_iso_mono": [2.7 * 1.380649e-16 * 2.72548 / 8.187105776823886e-07 * 2, 7.5657e-15 * 2.72548**4 * 16],
               "ec_ps_behind_jet": [cases.EPS_LINE, 0.024 * 2e46, 1e19],
               "tau_blr": [cases.L_DISK, 0.024, cases.EPS_LINE, 1e17, 10 ** rng.uniform(16, 17)],
               "tau_dt": [cases.L_DISK, 0.1, cases.EPS_DT, cases.R_DT, 10 ** rng.uniform(16, 18)],
               "tau_ss_disk": [cases.M_BH, cases.L_DISK, 1 / 12, 6 * cases.R_G, 200 * cases.R_G, 10 ** rng.uniform(16, 17)],
               }.get(process, [])
        delta = rng.uniform(10, 40)
        if process in ("synch", "ssc"):
            ne = [1e-8, 1.9, 2.6, 1e4, 10.0, 1e6, 0, 0]
        else:
            ne = [2.5e-9, 2.0, 3.5, 1e4, 20.0, 5e7, 0, 0]
        _core.model_row(models, i, z=1.0 if process not in ("synch", "ssc") else 0.069,
                        d_L=cases.D_L_EC, delta_D=delta, mu_s=1.0 if process.startswith("tau") else b["mu_s"],
                        B=10 ** rng.uniform(-1.5, 0), R_b=1e16, ne=ne, tgt=tgt)
    return models

def run(label, process, n_models, pts_per_sed, flops_pt, reps=5, **kw):
    plan = SedBatch(process, kw.pop("nu"), "BrokenPowerLaw", **kw)
    models = models_for(process, n_models)
    d_models = torch.from_numpy(np.ascontiguousarray(models).view(np.float64).reshape(n_models, -1).copy()).cuda()
    d_out = torch.empty(n_models, plan.n_nu, dtype=torch.float64, device="cuda")
    for _ in range(2): plan.launch(d_models, d_out)
    torch.cuda.synchronize()
    _lib.kernel_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): plan.launch(d_models, d_out)
    e1.record(); torch.cuda.synchronize()
    kms, kn = _lib.kernel_timing_read(); _lib.kernel_timing(False)
    ms = e0.elapsed_time(e1) / reps
    k_ms = kms / max(kn, 1)
    pts = pts_per_sed * n_models
    row = dict(case=label, models=n_models, ms_per_call=ms, main_kernel_ms=k_ms, sed_per_s=n_models / ms * 1e3,
               gpts_per_s=pts / ms / 1e6, nominal_tflops=pts * flops_pt / (k_ms * 1e-3) / 1e12,
               nominal_frac=pts * flops_pt / (k_ms * 1e-3) / peak, finite=bool(torch.isfinite(d_out).all()))
    rows.append(row)
    print(json.dumps(row), flush=True)

c1 = cases.case_c1()
g2 = bench.grids()
tnu = cases.case_tau()["nu"]
for n in (1, 256):
    run("C1 synch 100nu x 200g", "synch", n, 100 * 200, 24, nu=c1["nu"], gamma=c1["gamma"])
    run("C1 synch+ssa loglog", "synch", n, 100 * 200, 38, nu=c1["nu"], gamma=c1["gamma"], ssa=True, integrator="trapz_loglog")
    run("C1 ssc 100nu x 200eps x 200g", "ssc", n, 100 * 200 * 200 + 200 * 200, 24, nu=c1["nu"], gamma=c1["gamma"])
    run("C1 ssc loglog", "ssc", n, 100 * 200 * 200 + 200 * 200, 38, nu=c1["nu"], gamma=c1["gamma"], integrator="trapz_loglog")
    run("C2 ec_blr 200nu", "ec_blr", n, 2e8, 16, nu=g2["nu"], gamma=g2["gamma"], mu=g2["mu"], phi=g2["phi"])
    run("C2 ec_blr trapz_prefix", "ec_blr", n, 2e8, 16, nu=g2["nu"], gamma=g2["gamma"], mu=g2["mu"], phi=g2["phi"], integrator="trapz_prefix")
    run("C2 ec_blr loglog", "ec_blr", min(n, 16), 2e8, 30, nu=g2["nu"], gamma=g2["gamma"], mu=g2["mu"], phi=g2["phi"], integrator="trapz_loglog")
    run("C3 ec_dt 200nu", "ec_dt", n, 200 * 200 * 50, 16, nu=g2["nu"], gamma=g2["gamma"], phi=g2["phi"])
    run("ec_ss_disk 200nu", "ec_ss_disk", min(n, 64), 2e8, 16, nu=g2["nu"], gamma=g2["gamma"], phi=g2["phi"])
    run("ec_iso_mono 200nu", "ec_iso_mono", min(n, 64), 2e8, 16, nu=g2["nu"], gamma=g2["gamma"], mu=g2["mu"], phi=g2["phi"])
    run("ec_ps_behind_jet 200nu", "ec_ps_behind_jet", n, 200 * 200, 16, nu=g2["nu"], gamma=g2["gamma"])
    run("C3 tau_blr 200nu", "tau_blr", min(n, 32), 200 * 100 * 50 * 50, 21, nu=tnu)
    run("C3 tau_dt 200nu", "tau_dt", n, 200 * 50 * 50, 21, nu=tnu)
    run("tau_ss_disk 200nu", "tau_ss_disk", min(n, 32), 200 * 100 * 50 * 50, 21, nu=tnu)
Path("gpurun_out").mkdir(exist_ok=True)
json.dump(dict(fp64_peak_tflops=peak / 1e12, rows=rows), open("gpurun_out/time_all.json", "w"), indent=1)
