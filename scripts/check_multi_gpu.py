"""torchrun --nproc-per-node N scripts/check_multi_gpu.py : model sharding (config 5 style),
frequency sharding (config 4 style) and FitSedBatch over NCCL must reproduce the single-rank
result bit for bit; prints one JSON line from rank 0."""
import json, os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch, torch.distributed as dist
import bench
from agnpy_b200.batch import SedBatch
from agnpy_b200.fit import FitSedBatch
from tests import cases

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
saved = os.dup(1); os.dup2(2, 1)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dist.all_reduce(torch.zeros(1, device="cuda")); torch.cuda.synchronize()
sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
out = {"world": world}
g = bench.grids()
models = bench.synthetic_models(8 * world + 3)          # not a multiple of world: padding path
kw = dict(gamma=g["gamma"], mu=g["mu"], phi=g["phi"])
single = SedBatch("ec_blr", g["nu"], "BrokenPowerLaw", **kw)._local(models).cpu().numpy()
shard_m = SedBatch("ec_blr", g["nu"], "BrokenPowerLaw", group=dist.group.WORLD, **kw)(models)
out["model_sharding_bit_identical"] = bool(np.array_equal(single, shard_m))
# the other ways of joining the shards (agnpy_b200.batch.GATHER_MODES)
sb = SedBatch("ec_blr", g["nu"], "BrokenPowerLaw", group=dist.group.WORLD, **kw)
host = sb(models, gather="host")
out["gather_host_bit_identical"] = bool(np.array_equal(single, host))
host2 = sb(models[::-1].copy(), gather="host", copy=False)      # second call: the other half of the block
out["gather_host_second_call"] = bool(np.array_equal(single[::-1], host2))
r0 = sb(models, gather="rank0")
out["gather_rank0"] = bool(np.array_equal(single, r0)) if rank == 0 else r0 is None
per = -(-len(models) // world)
mine = sb(models, gather="none")
out["gather_none_is_own_shard"] = bool(np.array_equal(single[rank * per:(rank + 1) * per], mine))
pre = SedBatch("ec_blr", g["nu"], "BrokenPowerLaw", integrator="trapz_prefix", group=dist.group.WORLD, **kw)
dev, zeros_ok = cases.max_rel_dev(pre(models, gather="host"), single)
out["prefix_sharded_rel_dev"] = float(dev)
# config 4 in miniature (1000 gamma, 200 mu, 100 phi would take seconds per rank: use 400 x 120 x 60)
c = cases.case_c2(n_nu=50 * world + 1, n_gamma=400, n_mu=120, n_phi=60, r=2e17)
kw4 = dict(gamma=c["gamma"], mu=c["mu"], phi=c["phi"])
one = bench.synthetic_models(1)
single4 = SedBatch("ec_blr", c["nu"], "BrokenPowerLaw", **kw4)._local(one).cpu().numpy()
t0 = time.perf_counter()
shard_nu = SedBatch("ec_blr", c["nu"], "BrokenPowerLaw", group=dist.group.WORLD, shard="nu", **kw4)(one)
out["nu_sharding_bit_identical"] = bool(np.array_equal(single4, shard_nu))
# fit-level batch
rng = np.random.default_rng(5)
B = 4 * world + 1
pars = np.zeros((B, 23))
for i in range(B):
    d = rng.uniform(10, 40)
    pars[i] = [rng.uniform(-9, -7), rng.uniform(1.5, 2.5), rng.uniform(3, 4), rng.uniform(3.5, 5), np.log10(20),
               np.log10(5e7), 1.0, d, rng.uniform(-1.5, 0), 86400.0, (1 - 1 / d**2) / np.sqrt(1 - 1 / d**2),
               rng.uniform(17, 19), np.log10(2e46), cases.M_BH, 1e26, 6 * cases.R_G, 200 * cases.R_G, 0.024,
               1215.67, 1e17, 0.1, 1e3, cases.R_DT]
nu = np.logspace(10, 28, 100)
ref = FitSedBatch(nu, scenario="ec_blr_dt")(pars, cases.D_L_EC)
got = FitSedBatch(nu, scenario="ec_blr_dt", group=dist.group.WORLD)(pars, cases.D_L_EC)
out["fit_batch_bit_identical"] = bool(np.array_equal(ref, got))
got_h = FitSedBatch(nu, scenario="ec_blr_dt", group=dist.group.WORLD)(pars, cases.D_L_EC, gather="host")
out["fit_batch_host_gather_bit_identical"] = bool(np.array_equal(ref, got_h))
out["fit_batch_finite_positive"] = bool(np.all(np.isfinite(got)) and got.max() > 0)
out["prefix_sharded_ok"] = bool(zeros_ok and dev < 1e-12)
flags = torch.tensor([float(all(v for k, v in out.items() if k not in ("world", "prefix_sharded_rel_dev")))],
                     device="cuda")
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
out["all_ranks_ok"] = bool(flags.item() == 1.0)
if rank == 0:
    print(json.dumps(out), flush=True)
dist.barrier(); dist.destroy_process_group()
