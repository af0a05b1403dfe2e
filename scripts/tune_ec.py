"""Development aid: time ec_main variants (AGNPY_B200_EC_PLAN / AGNPY_B200_EC_NU_CHUNK knobs)."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bench
from agnpy_b200 import _lib
from agnpy_b200.batch import SedBatch

torch.cuda.set_device(0)
g = bench.grids()
def run(n_models, reps=8):
    models = bench.synthetic_models(n_models)
    plan = SedBatch("ec_blr", g["nu"], "BrokenPowerLaw", gamma=g["gamma"], mu=g["mu"], phi=g["phi"])
    d_models = torch.from_numpy(np.ascontiguousarray(models).view(np.float64).reshape(n_models, -1).copy()).cuda()
    d_out = torch.empty(n_models, g["nu"].size, dtype=torch.float64, device="cuda")
    for _ in range(3): plan.launch(d_models, d_out)
    torch.cuda.synchronize()
    _lib.kernel_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): plan.launch(d_models, d_out)
    e1.record(); torch.cuda.synchronize()
    kms, kn = _lib.kernel_timing_read(); _lib.kernel_timing(False)
    return e0.elapsed_time(e1) / reps, kms / max(kn, 1), d_out.cpu().numpy()

ref = None
variants = [(None, p, sg) for p in (None,) for sg in (None, "4")]
for v4, plan_s, sg in variants:
    if True:
        if v4: os.environ["AGNPY_B200_EC_V4"] = v4
        else: os.environ.pop("AGNPY_B200_EC_V4", None)
        if plan_s: os.environ["AGNPY_B200_EC_PLAN"] = plan_s
        else: os.environ.pop("AGNPY_B200_EC_PLAN", None)
        if sg: os.environ["AGNPY_B200_EC_LNU"] = sg
        else: os.environ.pop("AGNPY_B200_EC_LNU", None)
        res = []
        for n in (1, 64, 256):
            ms, kms, out = run(n)
            if ref is None and n == 64: ref = out
            if n >= 64: dev = float(np.nanmax(np.abs(out[:64] / np.where(ref != 0, ref, 1) - 1) * (ref != 0)))
            res.append(f"n={n}: {ms*1e3:8.1f} us (main {kms*1e3:8.1f} us)")
        print(f"v4={v4} plan={plan_s} sg={sg}: " + " | ".join(res) + f" | dev vs first {dev:.1e}", flush=True)
