"""Short EC-BLR run for ncu: a few batched launches of the C2 workload (8 models)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bench
from agnpy_b200.batch import SedBatch

torch.cuda.set_device(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
g = bench.grids()
models = bench.synthetic_models(n)
plan = SedBatch("ec_blr", g["nu"], "BrokenPowerLaw", gamma=g["gamma"], mu=g["mu"], phi=g["phi"])
for _ in range(3):
    out = plan(models)
torch.cuda.synchronize()
print("ok", out.shape, float(out.max()))
