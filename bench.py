#!/usr/bin/env python
"""bench.py -- SED evaluations/s of the emission-integral hot path (BASELINE.json metric).

Workload (BASELINE.json configs[1], SURVEY.md 8(d) "C2"): external Compton on the Lyman-alpha
spherical-shell BLR, default grids 200 nu x 200 gamma x 100 mu x 50 phi = 2e8 integrand points
per SED.  A *step* evaluates a batch of MODELS_PER_GPU independent parameter vectors per GPU
(different r, delta_D, electron spectrum...) through the batched entry point; at N>1 the batch
is sharded over ranks (weak scaling) and the SED blocks are joined by one NCCL all-gather
inside the timed step.

  python bench.py [--gpus N] [--steps K] [--warmup W]          # CUDA arm
  python bench.py --impl reference [...]                       # CPU arm: the unmodified reference, all cores

One JSON line on stdout (rank 0).  `value` is device-resident throughput (CUDA events, max over
ranks); `e2e` goes through the public host-buffer API (pinned H2D of the model table, D2H of
the SEDs, wall clock); `roofline` is for the dominant kernel (ec_main_poly_kernel): `frac` is the
EXECUTED FP64 work over the measured DFMA peak, the nominal 16 flop/point figure of SURVEY 8(a) is
kept beside it as `algorithmic`; `cpu_baseline` is the reference's own code (agnpy v0.5.0 from
baseline/_ref, on the miniature astropy of oracle/astropy_stub) on one core of this box, or the
oracle port where the reference install did not travel.  `extra` carries, at every N, the second
half of BASELINE.json's metric -- SSC SED evaluations/s at configs[0], measured as a weak-scaling step
like the headline -- and the two strong-scaling workloads: configs[4] (65 536 fit parameter vectors
in total) and configs[3] (one 2e10-point SED, frequency-sharded).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

MODELS_PER_GPU = 256
N_NU, N_GAMMA, N_MU, N_PHI = 200, 200, 100, 50
FLOPS_PER_POINT = 16          # nominal FP64 ops per compton_kernel point (np.trapz)
EXEC_FLOPS_PER_POINT = 4      # executed: 2 DFMA per visited point (polynomial form, FMA = 2)
# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, one launch of 256 models, from the
# ncu --set full capture summarised in profiles/r2_final_kernels_full.txt (reported as roofline.traffic)
NCU_DRAM_BYTES_PER_LAUNCH = 358936576
METRIC = "SED evals/sec (EC-BLR)"
UNIT = "SED/s"
WORKLOAD = ("configs[1]: ExternalCompton on SphericalShellBLR (Lyalpha), 200 nu x 200 gamma x "
            "100 mu x 50 phi, np.trapz")


# ---------------------------------------------------------------------------------------------
def synthetic_models(n, seed=20261019):
    """parameter vectors in the fit ranges of SURVEY.md 8(d) C5 (agnpy/fit/core.py:92-157) on the
    C2 physics: returns a numpy array of agnpy_b200._lib.MODEL_DTYPE"""
    from agnpy_b200 import _core, _lib
    from tests import cases
    rng = np.random.default_rng(seed)
    models = _lib.make_models(n)
    for i in range(n):
        delta = rng.uniform(10, 40)
        beta = np.sqrt(1 - 1 / delta**2)
        mu_s = (1 - 1 / (delta * delta)) / beta
        ne = [10 ** rng.uniform(-9, -7), rng.uniform(1.5, 2.5), rng.uniform(3, 4),
              10 ** rng.uniform(3.5, 5), 20.0, 5e7, 0.0, 0.0]
        R_b = 2.99792458e10 * 86400.0 * delta / (1 + cases.Z_EC)
        _core.model_row(models, i, z=cases.Z_EC, d_L=cases.D_L_EC, delta_D=delta, mu_s=mu_s,
                        B=10 ** rng.uniform(-1.5, 0), R_b=R_b, ne=ne,
                        tgt=[cases.L_DISK, cases.XI_LINE, cases.EPS_LINE, cases.R_LINE,
                             10 ** rng.uniform(17, 19)])
    return models


def grids():
    return dict(nu=np.logspace(15, 30, N_NU), gamma=np.logspace(1, 9, N_GAMMA),
                mu=np.linspace(-1, 1, N_MU), phi=np.linspace(0, 2 * np.pi, N_PHI))


def oracle_sed(model, g, nu=None):
    from oracle import agnpy_oracle as orc
    nu = g["nu"] if nu is None else nu
    return orc.ec_blr_sed_flux(nu, model["z"], model["d_L"], model["delta_D"], model["mu_s"],
                               model["R_b"], model["tgt"][0], model["tgt"][1], model["tgt"][2],
                               model["tgt"][3], model["tgt"][4], "BrokenPowerLaw",
                               tuple(model["ne"][:6]), gamma=g["gamma"], mu=g["mu"], phi=g["phi"])


def reference_module():
    """the unmodified reference (agnpy v0.5.0) if its offline install travelled with the repo
    (baseline/_ref) or its checkout is present; None otherwise"""
    try:
        from oracle import reference_loader
        if not reference_loader.available():
            return None
        return reference_loader.load()
    except Exception:
        return None


def reference_sed(agnpy, model, g, nu=None):
    """ExternalCompton.evaluate_sed_flux_blr of the reference itself (external_compton.py:398-490),
    Quantities in and out"""
    import astropy.units as u
    nu = g["nu"] if nu is None else nu
    t = model["tgt"]
    ne = model["ne"]
    with np.errstate(all="ignore"):
        sed = agnpy.ExternalCompton.evaluate_sed_flux_blr(
            nu * u.Hz, float(model["z"]), float(model["d_L"]) * u.cm, float(model["delta_D"]),
            float(model["mu_s"]), float(model["R_b"]) * u.cm, float(t[0]) * u.Unit("erg s-1"),
            float(t[1]), float(t[2]), float(t[3]) * u.cm, float(t[4]) * u.cm, agnpy.BrokenPowerLaw,
            float(ne[0]) * u.Unit("cm-3"), *[float(v) for v in ne[1:6]], integrator=np.trapz,
            gamma=g["gamma"].copy(), mu=g["mu"], phi=g["phi"])
    return np.asarray(sed.value, dtype=np.float64)


def cpu_sed_chunked(agnpy, model, g, chunk=25):
    """one SED on one core, frequency-chunked so the full-grid temporaries stay bounded"""
    fn = (lambda nn: reference_sed(agnpy, model, g, nn)) if agnpy else (lambda nn: oracle_sed(model, g, nn))
    return np.concatenate([fn(g["nu"][i:i + chunk]) for i in range(0, N_NU, chunk)])


def executed_point_fraction(models, g):
    """fraction of the full (gamma, mu, phi, nu) grid the kernel actually visits: points with
    gamma >= gamma_min(nu, mu, phi) inside the non-zero range of n_e (the rest is exactly 0 in
    the reference too).  Reported next to the algorithmic figure, never folded into it."""
    from oracle import agnpy_oracle as orc
    tot = vis = 0
    for m in models[: min(len(models), 16)]:
        eps_s = orc.nu_to_epsilon_prime(g["nu"], m["z"])
        R_line, r = m["tgt"][3], m["tgt"][4]
        with np.errstate(all="ignore"):
            ms = orc.mu_star_shell(g["mu"], R_line, r)
            gmin = orc.gamma_min_compton(eps_s[None, None, :], m["tgt"][2], m["mu_s"],
                                         ms[:, None, None], g["phi"][None, :, None])
        gam = g["gamma"]
        live = (gam / m["delta_D"] >= m["ne"][4]) & (gam / m["delta_D"] <= m["ne"][5])
        klo, khi = np.argmax(live), len(gam) - 1 - np.argmax(live[::-1])
        k0 = np.searchsorted(gam, np.nan_to_num(gmin, nan=np.inf), side="left")
        vis += np.clip(khi + 1 - np.maximum(k0, klo), 0, None).sum()
        tot += gmin.size * len(gam)
    return float(vis) / float(tot)


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None
        self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "50", "-i", str(self.index)], stdout=self.file, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.file.flush()
        rows = [r.split(",") for r in Path(self.file.name).read_text().splitlines() if r.strip()]
        os.unlink(self.file.name)
        sm, reasons = [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                out["sm_max_mhz"] = float(r[2])
                for nm, v in zip(names, r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


# ---------------------------------------------------------------------------------------------
C5_VECTORS = 65536


def _max_over_ranks(torch, dist, dev, world, x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def bench_c5(torch, dist, dev, world, group, barrier):
    """BASELINE configs[4]: 65 536 fit parameter vectors IN TOTAL (Synch + SSC + EC-BLR + EC-DT + the
    two thermal terms, sherpa-wrapper parameter order and grids, 100 nu), sharded over the ranks, one
    all-gather, result on the host of every rank.  Strong scaling: the total is fixed."""
    from agnpy_b200.fit import FitSedBatch
    from tests import cases
    pars = cases.c5_parameter_vectors(C5_VECTORS)
    out = {"vectors_total": C5_VECTORS, "n_gpus": world, "scaling": "strong",
           "workload": "configs[4]: Synch + SSC + EC-BLR + EC-DT + disk / torus black bodies per fit "
                       "vector (sherpa-wrapper order, gamma 200 / 300, 100 mu x 50 phi, 100 nu); d_L "
                       "from z on the host (Planck18, cached)",
           "points_per_vector": 100 * 200 + 200 * 200 + 100 * 200 * 200 + 100 * 300 * 5000 + 100 * 300 * 50}
    for integ in ("trapz", "trapz_prefix"):
        model = FitSedBatch(cases.C5_NU, "BrokenPowerLaw", scenario="ec_blr_dt", integrator=integ,
                            device=dev, group=group)
        gather = "host" if world > 1 else "all"
        model(pars, gather=gather, copy=False)          # warm-up: buffers, scratch, shared host block
        barrier()
        t0 = time.perf_counter()
        sed = model(pars, gather=gather, copy=False)
        torch.cuda.synchronize()
        dt = _max_over_ranks(torch, dist, dev, world, time.perf_counter() - t0)
        ok = bool(np.isfinite(sed).all() and (sed > 0).any() and sed.shape == (C5_VECTORS, 100))
        out[integ] = {"seconds": dt, "model_seds_per_s": C5_VECTORS / dt, "finite": ok,
                      "integrand_gpts_per_s": C5_VECTORS * out["points_per_vector"] / dt / 1e9}
        if getattr(model, "_host_gather", None) is not None:
            barrier()
            model._host_gather.close()
        del model
    out["note"] = ("end to end through agnpy_b200.fit.FitSedBatch.__call__: host parameter table in, "
                   "[65536, 100] SEDs in (node-shared, for N > 1) pinned host memory out on every rank; 'trapz' = the direct "
                   "per-point kernels, 'trapz_prefix' = the same np.trapz integral with the (mu, phi) "
                   "sum taken first (DESIGN.md section 4)")
    return out


def ssc_models_for(rank):
    """256 synthetic SSC models of this rank: the tutorial blob of BASELINE configs[0] with randomised
    Doppler factor, field, size and electron spectrum"""
    from agnpy_b200 import _core, _lib
    from tests import cases
    c1 = cases.case_c1()
    models = _lib.make_models(MODELS_PER_GPU)
    rng = np.random.default_rng(1 + rank)
    for i in range(MODELS_PER_GPU):
        _core.model_row(models, i, z=c1["z"], d_L=c1["d_L"], delta_D=rng.uniform(5, 30),
                        B=10 ** rng.uniform(-1.5, 0.5), R_b=10 ** rng.uniform(15.5, 16.5),
                        ne=[10 ** rng.uniform(-9, -7), rng.uniform(1.5, 2.5), rng.uniform(2.6, 4),
                            10 ** rng.uniform(3.5, 5), 10.0, 1e6, 0, 0])
    return c1, models


def bench_ssc(torch, dist, dev, world, rank, barrier, flush, steps):
    """the second half of BASELINE.json's metric: SSC SED evaluations/s (configs[0] grids), weak scaling
    like the headline -- 256 models per GPU per step, one all-gather of the SEDs inside the step, CUDA
    events, L2 flushed between steps, max over ranks"""
    from agnpy_b200.batch import SedBatch
    c1, models = ssc_models_for(rank)
    plan = SedBatch("ssc", c1["nu"], "BrokenPowerLaw", gamma=c1["gamma"], device=dev)
    d_models = torch.from_numpy(np.ascontiguousarray(models).view(np.float64)
                                .reshape(MODELS_PER_GPU, -1).copy()).to(dev)
    n_nu = c1["nu"].size
    d_out = torch.empty(MODELS_PER_GPU, n_nu, dtype=torch.float64, device=dev)
    gathered = torch.empty(MODELS_PER_GPU * world, n_nu, dtype=torch.float64, device=dev)

    def step():
        plan.launch(d_models, d_out)
        if world > 1:
            dist.all_gather_into_tensor(gathered, d_out)
    for _ in range(3):
        step()
    barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for e0, e1 in evs:
        flush.fill_(1.0)
        e0.record()
        step()
        e1.record()
    barrier()
    ms = _max_over_ranks(torch, dist, dev, world, sum(a.elapsed_time(b) for a, b in evs)) / steps
    pts = MODELS_PER_GPU * world * (n_nu * 200 * c1["gamma"].size)
    return {"value": MODELS_PER_GPU * world / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "scaling": "weak",
            "ms_per_step": ms, "integrand_gpts_per_s": pts / (ms * 1e-3) / 1e9,
            "collective": "none" if world == 1 else "one NCCL all_gather_into_tensor of 256x100 f64 per rank"}


def bench_c4(torch, dist, dev, world, group, barrier):
    """BASELINE configs[3]: one EC-BLR SED on 1000 nu x 1000 gamma x 200 mu x 100 phi = 2e10 points,
    the frequency grid dealt round-robin over the ranks, one all-gather.  Device time per SED
    (CUDA events around the launches and the collective, max over ranks)."""
    from agnpy_b200.batch import SedBatch
    nu = np.logspace(15, 30, 1000)
    plan = SedBatch("ec_blr", nu, "BrokenPowerLaw", gamma=np.logspace(1, 9, 1000),
                    mu=np.linspace(-1, 1, 200), phi=np.linspace(0, 2 * np.pi, 100), device=dev,
                    group=group, shard="nu" if world > 1 else "models")
    model = synthetic_models(1)
    for _ in range(3):
        out = plan.evaluate_device(model)
    barrier()
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        out = plan.evaluate_device(model)
    e1.record()
    torch.cuda.synchronize()
    wall = _max_over_ranks(torch, dist, dev, world, (time.perf_counter() - t0) / reps)
    ms = _max_over_ranks(torch, dist, dev, world, e0.elapsed_time(e1) / reps)
    sed = out.cpu().numpy()
    return {"n_gpus": world, "scaling": "strong", "points_per_sed": 2e10, "ms_per_sed": ms,
            "wall_ms_per_sed": wall * 1e3, "integrand_gpts_per_s": 2e10 / (ms * 1e-3) / 1e9,
            "finite": bool(np.isfinite(sed).all() and (sed > 0).sum() > 500),
            "workload": "configs[3]: EC-BLR, 1000 nu x 1000 gamma x 200 mu x 100 phi, np.trapz, "
                        "nu dealt round-robin over the ranks + one all-gather",
            "limits_at_large_n": "per-rank work is 4 launches + 1 all-gather of 1 KB around "
                                 "~0.1 ms of kernels: launch and collective latency, not FP64"}


# ---------------------------------------------------------------------------------------------
def run_cuda(args):
    import torch
    import torch.distributed as dist
    from agnpy_b200 import _lib
    from agnpy_b200.batch import SedBatch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (agnpy_b200 has no CPU fallback; "
                         "use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout from C; keep stdout to the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    n_gpus = world
    if args.gpus != world and rank == 0:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)

    g = grids()
    total = MODELS_PER_GPU * world
    models = synthetic_models(total)
    plan = SedBatch("ec_blr", g["nu"], "BrokenPowerLaw", gamma=g["gamma"], mu=g["mu"], phi=g["phi"],
                    device=dev)
    lo, hi = rank * MODELS_PER_GPU, (rank + 1) * MODELS_PER_GPU
    local = models[lo:hi]
    d_models = torch.from_numpy(np.ascontiguousarray(local).view(np.float64)
                                .reshape(len(local), -1).copy()).to(dev)
    d_out = torch.empty(len(local), N_NU, dtype=torch.float64, device=dev)
    gathered = torch.empty(total, N_NU, dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)  # > 126 MB L2

    def step():
        plan.launch(d_models, d_out)
        if world > 1:
            dist.all_gather_into_tensor(gathered, d_out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident, CUDA events per step, L2 flushed between steps --------------
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(args.steps)]
    barrier()
    for e0, e1 in evs:
        flush.fill_(1.0)
        e0.record()
        step()
        e1.record()
    barrier()
    launches = _lib.launch_count() - launches0
    ms_total = sum(e0.elapsed_time(e1) for e0, e1 in evs)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = total / (ms_per_step * 1e-3)

    # ---- e2e: public host-buffer API (pinned H2D of models, kernels, gather, D2H), wall clock ---
    plan_e2e = SedBatch("ec_blr", g["nu"], "BrokenPowerLaw", gamma=g["gamma"], mu=g["mu"],
                        phi=g["phi"], device=dev, group=None if world == 1 else dist.group.WORLD)
    def e2e_rate(gather):
        for _ in range(3):
            r = plan_e2e(models, gather=gather, copy=False)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            r = plan_e2e(models, gather=gather, copy=False)
        torch.cuda.synchronize()
        dt = _max_over_ranks(torch, dist, dev, world, time.perf_counter() - t0)
        return total * args.steps / dt, r
    # every rank ends up with the whole [N*256, 200] batch in host memory.  N > 1: assembled in
    # node-shared pinned host memory (each rank copies only its own shard device->host; no device
    # collective on this path); the variant through the NCCL all-gather + full D2H per rank is
    # reported beside it.
    e2e_gather = "host" if world > 1 else "all"
    e2e_value, res = e2e_rate(e2e_gather)
    e2e_allgather_value = e2e_rate("all")[0] if world > 1 else None
    res = np.array(res[:1])                      # row 0 for the parity spot check below
    clocks = sampler.stop() if rank == 0 else None

    # ---- strong-scaling workloads of BASELINE.json, at every N (all ranks take part) ------------
    group = None if world == 1 else dist.group.WORLD
    scaling = {"c5_fit_batch": bench_c5(torch, dist, dev, world, group, barrier),
               "c4_nu_sharded": bench_c4(torch, dist, dev, world, group, barrier)}
    ssc_scaling = bench_ssc(torch, dist, dev, world, rank, barrier, flush, max(3, min(args.steps, 10)))

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- rank 0 only from here: parity spot check, roofline, single-call latency, CPU baseline --
    from tests import cases
    chk = oracle_sed(models[0], g, nu=g["nu"][::8])
    dev_rel, zeros_ok = cases.max_rel_dev(res[0][::8], chk)
    agnpy_ref = reference_module()
    ref_dev = None
    if agnpy_ref is not None:       # the same frequencies through the reference's own code
        ref_dev, _ = cases.max_rel_dev(res[0][::8], reference_sed(agnpy_ref, models[0], g, g["nu"][::8]))

    peak = _lib.fp64_peak(300.0)
    _lib.kernel_timing(True)
    for _ in range(max(3, min(args.steps, 10))):
        plan.launch(d_models, d_out)
    torch.cuda.synchronize()
    kms, kn = _lib.kernel_timing_read()
    _lib.kernel_timing(False)
    k_ms = kms / max(kn, 1)
    pts_per_launch = len(local) * N_NU * N_GAMMA * N_MU * N_PHI
    achieved = pts_per_launch * FLOPS_PER_POINT / (k_ms * 1e-3) / 1e12
    frac_pts = executed_point_fraction(local, g)
    # mirror-symmetric phi grid: each mirror pair of (mu, phi) points is evaluated once
    folded = bool(plan.integ & _lib.INTEG_FOLD_PHI) and not os.environ.get("AGNPY_B200_EC_NOFOLD")
    if folded:
        frac_pts *= ((N_PHI + 1) // 2) / N_PHI
    executed = pts_per_launch * frac_pts * EXEC_FLOPS_PER_POINT / (k_ms * 1e-3) / 1e12
    # single-call drop-in latency (the literal reference call with Quantities)
    import agnpy_b200 as ab
    from agnpy_b200 import spectra, u
    c = cases.case_c2()
    ne = spectra.BrokenPowerLaw(*c["ne_args"])
    qa = (c["ne_args"][0] * u.Unit("cm-3"),) + tuple(c["ne_args"][1:])

    def one():
        return ab.ExternalCompton.evaluate_sed_flux_blr(
            c["nu"] * u.Hz, c["z"], c["d_L"] * u.cm, c["delta_D"], c["mu_s"], c["R_b"] * u.cm,
            c["L_disk"] * u.Unit("erg s-1"), c["xi_line"], c["epsilon_line"], c["R_line"] * u.cm,
            c["r"] * u.cm, ne, *qa, gamma=c["gamma"], mu=c["mu"], phi=c["phi"])
    for _ in range(5):
        one()
    t0 = time.perf_counter()
    for _ in range(50):
        one()
    single_us = (time.perf_counter() - t0) / 50 * 1e6

    # ---- secondary figures (same run, device-resident, CUDA events) ---------------------------
    def timed(plan_x, dm, dout, reps):
        for _ in range(3):
            plan_x.launch(dm, dout)
        torch.cuda.synchronize()
        _lib.kernel_timing(True)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            plan_x.launch(dm, dout)
        b.record()
        torch.cuda.synchronize()
        km, kk = _lib.kernel_timing_read()
        _lib.kernel_timing(False)
        return a.elapsed_time(b) / reps, km / max(kk, 1)

    # (a) the opt-in suffix-sum evaluation of np.trapz on the same batch
    plan_fast = SedBatch("ec_blr", g["nu"], "BrokenPowerLaw", gamma=g["gamma"], mu=g["mu"],
                         phi=g["phi"], integrator="trapz_prefix", device=dev)
    fast_ms, fast_kms = timed(plan_fast, d_models, d_out, max(3, min(args.steps, 10)))
    fast_dev, _ = cases.max_rel_dev(d_out[0].cpu().numpy()[::8], chk)
    # (b) SSC at BASELINE configs[0] grids (tutorial blob, 100 nu x 200 seeds x 200 gamma)
    c1, ssc_models = ssc_models_for(0)
    plan_ssc = SedBatch("ssc", c1["nu"], "BrokenPowerLaw", gamma=c1["gamma"], device=dev)
    d_ssc = torch.from_numpy(np.ascontiguousarray(ssc_models).view(np.float64)
                             .reshape(MODELS_PER_GPU, -1).copy()).to(dev)
    d_ssc_out = torch.empty(MODELS_PER_GPU, c1["nu"].size, dtype=torch.float64, device=dev)
    ssc_ms, ssc_kms = timed(plan_ssc, d_ssc, d_ssc_out, max(3, min(args.steps, 10)))
    from oracle import agnpy_oracle as orc
    m0 = ssc_models[0]
    ssc_ref = orc.ssc_sed_flux(c1["nu"], m0["z"], m0["d_L"], m0["delta_D"], m0["B"], m0["R_b"],
                               "BrokenPowerLaw", tuple(m0["ne"][:6]), gamma=c1["gamma"])
    ssc_dev, _ = cases.max_rel_dev(d_ssc_out[0].cpu().numpy(), ssc_ref)
    ssc_pts = MODELS_PER_GPU * (100 * 200 * 200)
    extra = {
        "ec_blr_trapz_prefix": {
            "value": len(local) / (fast_ms * 1e-3), "unit": UNIT, "ms_per_step": fast_ms,
            "kernel": "ec_moment_kernel", "kernel_ms": fast_kms, "parity_max_rel_dev": fast_dev,
            "note": "opt-in integrator=trapz_prefix: the same np.trapz integral with the (mu, phi) sum "
                    "taken before the gamma sum (prefix moments of the sorted pair table), same grid "
                    "and masks (DESIGN.md section 4); not the headline"},
        "ssc_config1": {
            **ssc_scaling,
            "workload": "configs[0]: SSC of the tutorial blob, 100 nu x 200 seed nu x 200 gamma, "
                        "np.trapz, 256 models per GPU per step (seed synchrotron spectrum included)",
            "kernel": "ssc_trapz_kernel", "kernel_ms": ssc_kms,
            "one_gpu_no_flush": {"value": MODELS_PER_GPU / (ssc_ms * 1e-3), "ms_per_step": ssc_ms},
            "ncu_fp64_pipe_active_pct": 58.4,   # profiles/r2_final_kernels_full.txt, this batch size
            "roofline_frac_algorithmic": ssc_pts * 24 / (ssc_kms * 1e-3) / peak,
            "algorithmic_flops_per_point": 24, "parity_max_rel_dev": ssc_dev},
    }

    extra.update(scaling)

    # CPU baseline: the reference's own code (or the oracle port), 1 core, bounded sample
    t0 = time.perf_counter()
    n_cpu = 0
    while True:
        cpu_sed_chunked(agnpy_ref, models[n_cpu], g)
        n_cpu += 1
        if time.perf_counter() - t0 > 12 or n_cpu >= 3:
            break
    cpu_dt = time.perf_counter() - t0
    cpu_value = n_cpu / cpu_dt
    cpu_kind = "reference (agnpy 0.5.0 from baseline/_ref, CGS-stubbed astropy units)" if agnpy_ref \
        else "port"

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "models_per_gpu_per_step": MODELS_PER_GPU,
                   "points_per_sed": N_NU * N_GAMMA * N_MU * N_PHI,
                   "l2": "flushed between timed steps (256 MiB write); inputs are KB-sized, "
                         "the path is FP64-pipe bound",
                   "collective": "none" if world == 1 else "one NCCL all_gather_into_tensor of "
                                 f"{MODELS_PER_GPU}x{N_NU} f64 per rank inside the step",
                   "parity_vs_oracle_max_rel_dev": dev_rel, "parity_zeros_exact": zeros_ok,
                   "parity_vs_reference_max_rel_dev": ref_dev},
        "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": UNIT,
                "h2d_bytes_per_step": int(total * 176),
                "d2h_bytes_per_step": int(total * N_NU * 8),
                "api": f"agnpy_b200.batch.SedBatch.__call__(models, gather='{e2e_gather}', copy=False): host "
                       "numpy in, the whole [N*256, 200] block in pinned host memory out on every rank",
                "d2h_bytes_per_rank_per_step": int(MODELS_PER_GPU * N_NU * 8) if world > 1
                else int(total * N_NU * 8),
                "value_via_device_allgather": e2e_allgather_value,
                "single_call_us": single_us,
                "single_call_api": "ExternalCompton.evaluate_sed_flux_blr (Quantity in/out, "
                                   "agnpy_b200_ec_sed_host)"},
        "integrand_gpts_per_s": total * N_NU * N_GAMMA * N_MU * N_PHI / (ms_per_step * 1e-3) / 1e9,
        "roofline": {
            "bound": "fp64", "kernel": "ec_main_poly_kernel<5,128> (128 threads x 5 sorted pairs, 128 registers, 4 blocks/SM)",
            "achieved": executed, "peak": peak / 1e12, "unit": "TFLOP/s",
            "frac": executed / (peak / 1e12), "traffic": NCU_DRAM_BYTES_PER_LAUNCH if MODELS_PER_GPU == 256 else None,
            "what": "EXECUTED FP64 work of the dominant kernel: 2 DFMA (4 flop) per visited integrand "
                    "point x visited points per launch / CUDA-event time of that launch, over the "
                    "measured DFMA-chain peak of this GPU",
            "fp64_flops_per_visited_point": EXEC_FLOPS_PER_POINT,
            "visited_point_fraction": frac_pts, "phi_axis_folded": folded,
            "kernel_ms": k_ms, "kernel_launches_timed": int(kn),
            "kernel_share_of_step": k_ms / ms_per_step,
            "peak_source": "live DFMA-chain microbenchmark agnpy_b200_fp64_peak on this GPU "
                           "(MEASURED_PEAKS.json has no FP64 entry; nominal 148 x 64 x 2 x 1.965 GHz "
                           "= 37.2 TFLOP/s)",
            "algorithmic": {
                "flops_per_point": FLOPS_PER_POINT, "tflops": achieved,
                "over_peak": achieved / (peak / 1e12), "speedup_over_executed": achieved / executed,
                "note": "SURVEY 8(a) nominal count (16 flop per point of the full 2e8-point grid, two "
                        "divisions included) over the same time: above peak because the divisions and "
                        "all (gamma,nu)-only / (mu,phi)-only factors are hoisted into tables (2 DFMA "
                        "per point: wf K = P + Q b^2 + R b), points the reference masks to exactly 0 "
                        "are skipped and mirror pairs of the symmetric phi grid share one evaluation"},
            "traffic_note": "bytes per launch, dram__bytes_read.sum + dram__bytes_write.sum of ONE ncu --set full "
                            "capture of this kernel at this batch size (profiles/r2_final_kernels_full.txt), "
                            "not measured in this run; algorithmic in/out is 0.45 MB, the rest is the 328 MB of "
                            "polynomial rows written by ec_tables2_kernel and read once per pair slice "
                            "(1.7 % of HBM peak: the kernel is FP64-pipe bound)"},
        "cpu_baseline": {"value": cpu_value, "unit": UNIT, "cores": 1, "kind": cpu_kind,
                         "sample": f"{n_cpu} SED(s) of the same workload, nu-chunked by 25, 1 thread "
                                   "(the reference is single-threaded)",
                         "host_cpus": os.cpu_count()},
        "clocks": clocks,
        "extra": extra,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REF = None


def _ref_worker(job):
    model_bytes, lo, hi = job
    from agnpy_b200 import _lib
    model = np.frombuffer(model_bytes, dtype=_lib.MODEL_DTYPE)[0]
    g = grids()
    if _REF is not None:
        return reference_sed(_REF, model, g, g["nu"][lo:hi])
    return oracle_sed(model, g, nu=g["nu"][lo:hi])


def run_reference(args):
    """CPU arm: the reference's own implementation of the path on all host cores.  The reference
    (agnpy v0.5.0) is installed offline into baseline/_ref (`pip install --no-deps --target`), which
    travels with the repo; astropy has no wheel, so it runs on the miniature astropy of
    oracle/astropy_stub (units with astropy's semantics; the reference's own test-suite passes on
    it).  One step = one SED of the same workload through ExternalCompton.evaluate_sed_flux_blr,
    its 200 frequencies split over a process pool (the reference itself is single-threaded).  If
    the install is absent the oracle port of the same numpy arithmetic is timed instead."""
    global _REF
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    _REF = reference_module()       # imported before the fork: the workers inherit it
    cores = max(1, min(os.cpu_count() or 1, 32))
    chunk = 4
    models = synthetic_models(max(args.steps + args.warmup, 1))
    jobs_for = lambda m: [(m.tobytes(), lo, min(lo + chunk, N_NU)) for lo in range(0, N_NU, chunk)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for i in range(args.warmup):
            pool.map(_ref_worker, jobs_for(models[i:i + 1]))
        t0 = time.perf_counter()
        for i in range(args.steps):
            pool.map(_ref_worker, jobs_for(models[args.warmup + i:args.warmup + i + 1]))
        dt = time.perf_counter() - t0
    value = args.steps / dt
    kind = "reference" if _REF is not None else "port"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": "1 SED per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{args.steps} step(s) x 1 SED, nu split in chunks of {chunk} "
                                   f"over {cores} processes",
                         "host_cpus": os.cpu_count(),
                         "how": ("unmodified agnpy 0.5.0 from baseline/_ref: ExternalCompton."
                                 "evaluate_sed_flux_blr with Quantities; astropy replaced by "
                                 "oracle/astropy_stub (CGS-consistent units, no wheel for astropy in this "
                                 "image)") if _REF is not None else
                                "reference install absent: oracle port of its numpy arithmetic "
                                "(oracle/agnpy_oracle.py)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = 3 if args.steps is None else args.steps
        args.warmup = 1 if args.warmup is None else args.warmup
        run_reference(args)
    else:
        args.steps = 20 if args.steps is None else args.steps
        args.warmup = 3 if args.warmup is None else args.warmup
        run_cuda(args)


if __name__ == "__main__":
    main()
