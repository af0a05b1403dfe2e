"""Development aid: time the EC-BLR np.trapz main kernel under the A/B knobs of ec.cu
(AGNPY_B200_EC_PLAN="T,J" with T,J in {128,5 128,4 64,8 128,8}, AGNPY_B200_EC_LNU, AGNPY_B200_EC_NOFOLD;
earlier sweeps, with knobs and tile shapes that have since been removed, are kept in
profiles/r2_tune_ec.txt).
Usage: python scripts/tune_ec.py ["PLAN=128,4" "PLAN=64,8 LNU=4" ...]   (default: a built-in sweep)
Prints, per variant: wall per step and main-kernel time for 256 models, SED/s, and the largest
deviation from the first variant (all variants must agree to rounding)."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bench
from agnpy_b200 import _lib
from agnpy_b200.batch import SedBatch

torch.cuda.set_device(0)
g = bench.grids()
KNOBS = ("PLAN", "LNU", "NOFOLD")


def run(n_models, reps=8):
    models = bench.synthetic_models(n_models)
    plan = SedBatch("ec_blr", g["nu"], "BrokenPowerLaw", gamma=g["gamma"], mu=g["mu"], phi=g["phi"])
    d_models = torch.from_numpy(np.ascontiguousarray(models).view(np.float64).reshape(n_models, -1).copy()).cuda()
    d_out = torch.empty(n_models, g["nu"].size, dtype=torch.float64, device="cuda")
    for _ in range(3):
        plan.launch(d_models, d_out)
    torch.cuda.synchronize()
    _lib.kernel_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        plan.launch(d_models, d_out)
    e1.record()
    torch.cuda.synchronize()
    kms, kn = _lib.kernel_timing_read()
    _lib.kernel_timing(False)
    return e0.elapsed_time(e1) / reps, kms / max(kn, 1), d_out.cpu().numpy()


variants = sys.argv[1:] or ["PLAN=128,5", "PLAN=128,4", "PLAN=64,8", "PLAN=128,8", "PLAN=128,5 LNU=4", "PLAN=128,5 NOFOLD=1"]
ref = None
for v in variants:
    for k in KNOBS:
        os.environ.pop("AGNPY_B200_EC_" + k, None)
    for item in v.split():
        k, val = item.split("=")
        os.environ["AGNPY_B200_EC_" + k] = val
    try:
        ms1, k1, _ = run(1)
        ms, kms, out = run(256)
    except Exception as exc:  # a variant that does not launch must not stop the sweep
        print(f"{v:32s} FAILED: {exc}", flush=True)
        continue
    if ref is None:
        ref = out
    nz = ref != 0
    dev = float(np.max(np.abs(out[nz] / ref[nz] - 1)))
    same_zeros = bool(np.array_equal(out == 0, ref == 0))
    print(f"{v:32s} n=1: {ms1*1e3:7.1f} us (main {k1*1e3:6.1f}) | n=256: {ms:7.3f} ms (main {kms:7.3f} ms) "
          f"{256 / ms * 1e3:9.0f} SED/s | dev vs first {dev:.1e} zeros {same_zeros}", flush=True)
