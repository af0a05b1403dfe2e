"""BASELINE config 5 on the GPUs of one box: B fit parameter vectors (Synch + SSC + EC-BLR + EC-DT
+ thermal terms, wrapper grids, 100 nu), sharded over the ranks.  Run plain (1 GPU) or under torchrun.
Prints one JSON line from rank 0."""
import json, os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import torch.distributed as dist
from agnpy_b200.fit import FitSedBatch
from tests import cases

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
group = None
if world > 1:
    saved = os.dup(1); os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dist.all_reduce(torch.zeros(1, device="cuda")); torch.cuda.synchronize()
    sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
    group = dist.group.WORLD
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096 * world
integ = sys.argv[2] if len(sys.argv) > 2 else "trapz"
rng = np.random.default_rng(20261019)                      # SURVEY.md 8(d) C5 ranges
pars = np.zeros((B, 23))
delta = rng.uniform(10, 40, B)
pars[:, 0], pars[:, 1], pars[:, 2] = rng.uniform(-9, -7, B), rng.uniform(1.5, 2.5, B), rng.uniform(3, 4, B)
pars[:, 3], pars[:, 4], pars[:, 5] = rng.uniform(3.5, 5, B), np.log10(20), np.log10(5e7)
pars[:, 6], pars[:, 7], pars[:, 8], pars[:, 9] = 1.0, delta, rng.uniform(-1.5, 0, B), 86400.0
pars[:, 10] = (1 - 1 / delta**2) / np.sqrt(1 - 1 / delta**2)
pars[:, 11], pars[:, 12] = rng.uniform(17, 19, B), np.log10(2e46)
pars[:, 13:] = [cases.M_BH, 1e26, 6 * cases.R_G, 200 * cases.R_G, 0.024, 1215.67, 1e17, 0.1, 1e3, cases.R_DT]
nu = np.logspace(10, 28, 100)
model = FitSedBatch(nu, "BrokenPowerLaw", scenario="ec_blr_dt", group=group, integrator=integ)
sed = model(pars[: 64 * world], cases.D_L_EC)              # warm-up
torch.cuda.synchronize()
if world > 1: dist.barrier()
t0 = time.perf_counter()
sed = model(pars, cases.D_L_EC)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
t = torch.tensor([dt], device="cuda", dtype=torch.float64)
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    pts = B * (100 * 200 + 100 * 200 * 200 + 200 * 200 + 100 * 300 * 5000 + 100 * 300 * 50)
    print(json.dumps({"config": "C5 fit-level batch", "n_gpus": world, "models": B, "integrator": integ,
                      "seconds": float(t.item()), "model_seds_per_s": B / float(t.item()),
                      "integrand_gpts_per_s": pts / float(t.item()) / 1e9,
                      "finite": bool(np.isfinite(sed).all()), "positive": bool((sed > 0).any()),
                      "extrapolated_65536_models_s": 65536 / (B / float(t.item()))}), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()

# optional: where the time of one rank goes (python scripts/time_c5.py 16384 trapz breakdown)
if len(sys.argv) > 3 and sys.argv[3] == "breakdown" and world == 1:
    from agnpy_b200.fit import build_models
    t0 = time.perf_counter()
    models = build_models(pars, None, "BrokenPowerLaw", "ec_blr_dt")
    host_s = time.perf_counter() - t0
    parts = {"host build_models (d_L from z)": host_s * 1e3}
    for name, plan in model.plans.items():
        plan._local(models[name]); torch.cuda.synchronize()   # warm-up at full size (buffers)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record(); plan._local(models[name]); e1.record(); torch.cuda.synchronize()
        parts[name] = e0.elapsed_time(e1)
        parts[name + " (wall)"] = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter(); out = model(pars, copy=False); parts["whole call, copy=False (wall)"] = (time.perf_counter() - t0) * 1e3
    print(json.dumps({"breakdown_ms": parts, "models": B}), flush=True)
