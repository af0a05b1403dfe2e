"""BASELINE config 4: high-resolution EC-BLR sweep (1000 nu x 1000 gamma x 200 mu x 100 phi = 2e10
integrand points for ONE model), frequency-sharded over the ranks (plain: 1 GPU; torchrun: N GPUs,
one all-gather of the SED).  Prints one JSON line from rank 0; parity of a 16-frequency subsample
against the oracle is part of tests/test_gpu_full_size.py::test_config4_properties."""
import json, os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import torch.distributed as dist
import bench
from agnpy_b200 import _lib
from agnpy_b200.batch import SedBatch
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
c = cases.case_c2(n_nu=1000, n_gamma=1000, n_mu=200, n_phi=100, r=2e17)
model = bench.synthetic_models(1)
plan = SedBatch("ec_blr", c["nu"], "BrokenPowerLaw", gamma=c["gamma"], mu=c["mu"], phi=c["phi"],
                group=group, shard="nu" if world > 1 else "models")
for _ in range(3):
    out = plan.evaluate_device(model)
torch.cuda.synchronize()
if world > 1: dist.barrier()
reps = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
_lib.kernel_timing(True)
e0.record()
for _ in range(reps):
    out = plan.evaluate_device(model)
e1.record(); torch.cuda.synchronize()
kms, kn = _lib.kernel_timing_read(); _lib.kernel_timing(False)
ms = e0.elapsed_time(e1) / reps
t = torch.tensor([ms], device="cuda", dtype=torch.float64)
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    ms = float(t.item())
    sed = out.cpu().numpy()
    print(json.dumps({"config": "C4 EC-BLR 1000 nu x 1000 gamma x 200 mu x 100 phi, 1 model", "n_gpus": world,
                      "ms_per_sed": ms, "main_kernel_ms_rank0": kms / max(kn, 1),
                      "integrand_gpts_per_s": 2e10 / ms / 1e6, "finite": bool(np.isfinite(sed).all()),
                      "nonzero": int(np.count_nonzero(sed))}), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
