"""Batched, multi-GPU entry point for fit / MCMC callers (new API).

The reference's fit wrappers (agnpy/fit/sherpa_wrapper.py:51-325, gammapy_wrapper.py:161-357)
evaluate one parameter vector per call.  Here a batch of parameter vectors ("models") is
evaluated in one launch sequence on shared integration grids; with torch.distributed
initialised the batch is sharded over the ranks (one process per GPU) and the per-rank SED
blocks are joined with ONE NCCL all-gather (every model is independent: no other exchange).

    plan = SedBatch("ec_blr", nu_hz, "BrokenPowerLaw", gamma=..., mu=..., phi=...)
    sed = plan(models)            # models: numpy structured array of _lib.MODEL_DTYPE, host
                                  # returns numpy [n_models, n_nu]  (erg cm-2 s-1, or tau)

Grids live on the device for the life of the plan; per call the model table is copied
host->device from pinned memory and the result device->host into pinned memory.
"""
import os

import numpy as np

from . import _lib
from ._lib import check, lib

_NE_KINDS = {"PowerLaw": _lib.NE_POWERLAW, "BrokenPowerLaw": _lib.NE_BROKEN_POWERLAW,
             "LogParabola": _lib.NE_LOG_PARABOLA, "ExpCutoffPowerLaw": _lib.NE_EXP_CUTOFF_POWERLAW,
             "ExpCutoffBrokenPowerLaw": _lib.NE_EXP_CUTOFF_BROKEN_POWERLAW}
_EC = {"ec_iso_mono": _lib.EC_ISO_MONO, "ec_ps_behind_jet": _lib.EC_PS_BEHIND_JET,
       "ec_ss_disk": _lib.EC_SS_DISK, "ec_blr": _lib.EC_BLR, "ec_dt": _lib.EC_DT}
_TAU = {"tau_ss_disk": _lib.TAU_SS_DISK, "tau_blr": _lib.TAU_BLR, "tau_dt": _lib.TAU_DT,
        "tau_ps_behind_blob": _lib.TAU_PS_BEHIND_BLOB,
        "tau_ps_behind_blob_mu_s": _lib.TAU_PS_BEHIND_BLOB_MU_S,
        "tau_blr_mu_s": _lib.TAU_BLR_MU_S, "tau_dt_mu_s": _lib.TAU_DT_MU_S}
_TAU_OFF_AXIS = ("tau_ps_behind_blob_mu_s", "tau_blr_mu_s", "tau_dt_mu_s")
_BB = {"bb_disk": _lib.BB_DISK, "bb_disk_norm": _lib.BB_DISK_NORM, "bb_dt": _lib.BB_DT,
       "bb_dt_norm": _lib.BB_DT_NORM}
PROCESSES = ("synch", "ssc") + tuple(_EC) + tuple(_TAU) + tuple(_BB)


# ---- sharding arithmetic (pure; covered by the gloo tests on CPU) ---------------------------
def partition(n_models, world_size):
    """models per rank (padded so every rank holds the same count) -> (per_rank, padded_total)"""
    per_rank = -(-n_models // world_size)
    return per_rank, per_rank * world_size


def local_slice(n_models, world_size, rank):
    """[lo, hi) of the real models owned by `rank`, and the padded per-rank count"""
    per_rank, _ = partition(n_models, world_size)
    lo = min(rank * per_rank, n_models)
    hi = min(lo + per_rank, n_models)
    return lo, hi, per_rank


GATHER_MODES = ("all", "rank0", "none", "host")


class GatherBuffers:
    """device staging for the one collective of a sharded call, allocated once per shape and kept by the
    plan (a fit loop calls with the same batch size thousands of times)"""

    def __init__(self):
        self._key, self.block, self.out = None, None, None

    def get(self, rows, cols, world, device, need_out):
        import torch
        key = (rows, cols, world, str(device))
        if key != self._key:
            self._key = key
            self.block = torch.zeros(rows, cols, dtype=torch.float64, device=device)
            self.out = None
        if need_out and self.out is None:
            self.out = torch.empty(world * rows, cols, dtype=torch.float64, device=device)
        return self.block, self.out


HOST_GATHER_TIMEOUT_S = float(os.environ.get("AGNPY_B200_HOST_GATHER_TIMEOUT_S", "600"))


class HostGather:
    """gather="host": the ranks of ONE node assemble the batch in host memory, not on the device.

    When the consumer of the SEDs is host code (a sampler, a likelihood), the device all-gather plus
    a device->host copy of the WHOLE batch on every rank is wasted motion: here every rank copies
    only its own shard device->host, straight into its rows of one block of shared memory (a file in
    /dev/shm) that all ranks map (registered with CUDA, so the copy is an asynchronous pinned transfer), and a
    sequence-number barrier in the same segment tells everyone when all shards have landed.  Bytes
    over PCIe per rank: those of its shard, as at N = 1; no collective on the data path.

    Two blocks are used alternately: the block of call k is rewritten at call k + 2, and no rank can
    enter call k + 2 before every rank has entered call k + 1 (the barrier), i.e. has finished with
    the views handed out by call k -- the same "valid until the next call" rule as the pinned block
    of a single rank."""

    def __init__(self, group, rows, cols, register=True):
        import torch
        import torch.distributed as dist
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.rows, self.cols, self.group = rows, cols, group
        hosts = [None] * self.world
        import socket
        dist.all_gather_object(hosts, socket.gethostname(), group=group)
        if len(set(hosts)) != 1:
            raise RuntimeError("gather='host' needs all ranks on one node; use gather='all'")
        block = rows * cols * 8
        size = 2 * block + 64 * self.world + 64
        # a file in /dev/shm mapped by every rank (plain mmap: the mapping lives as long as the numpy
        # views do, and the name is unlinked as soon as everybody has it open)
        import mmap
        import os
        path = [None]
        if self.rank == 0:
            path[0] = f"/dev/shm/agnpy_b200_{os.getpid()}_{id(self):x}"
            fd = os.open(path[0], os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
            os.ftruncate(fd, size)
        dist.broadcast_object_list(path, src=dist.get_global_rank(group, 0) if group is not None else 0,
                                   group=group)
        if self.rank != 0:
            fd = os.open(path[0], os.O_RDWR)
        self._mmap = mmap.mmap(fd, size)
        os.close(fd)
        self._path = path[0]
        buf = self._mmap
        self.blocks = [np.frombuffer(buf, dtype=np.float64, count=rows * cols, offset=b * block)
                       .reshape(rows, cols) for b in range(2)]
        # one cache line per rank: the sequence number of the last call whose shard has landed
        self.flags = np.frombuffer(buf, dtype=np.int64, count=8 * self.world, offset=2 * block)
        self.seq = 0
        self._registered = False
        if register and torch.cuda.is_available():
            ptr = self.blocks[0].ctypes.data
            rc = torch.cuda.cudart().cudaHostRegister(ptr, 2 * block, 0)
            self._registered = int(rc) == 0
            self._ptr = ptr
        self.tensors = [torch.from_numpy(b) for b in self.blocks]
        dist.barrier(group=group)       # everyone has mapped the file before rank 0 unlinks it
        if self.rank == 0:
            os.unlink(self._path)       # the mappings stay alive; the name is freed at once

    def block(self):
        """(torch view, numpy view) of the block of the call that is about to be made"""
        b = (self.seq + 1) & 1
        return self.tensors[b], self.blocks[b]

    def publish_and_wait(self):
        """this rank's shard is in place: announce it, wait for everybody else's"""
        self.seq += 1
        self.flags[8 * self.rank] = self.seq
        others = self.flags[::8]
        spins, t0 = 0, None
        while int(others.min()) < self.seq:
            spins += 1
            if spins & 0xffff == 0:      # a rank that died must not leave the others spinning for ever
                import time
                t0 = t0 or time.monotonic()
                if time.monotonic() - t0 > HOST_GATHER_TIMEOUT_S:
                    late = [r for r in range(self.world) if int(others[r]) < self.seq]
                    raise RuntimeError(f"gather='host': ranks {late} did not deliver their shard of call "
                                       f"{self.seq} within {HOST_GATHER_TIMEOUT_S} s")

    def close(self):
        try:
            if self._registered:
                import torch
                torch.cuda.cudart().cudaHostUnregister(self._ptr)
            self.tensors = self.blocks = self.flags = None   # the mapping goes with its last view
        except Exception:
            pass


def host_sharded_evaluate(local_compute, models, n_nu, hg, stream_sync=None):
    """gather="host" (see HostGather): returns the numpy view [n, n_nu] of the shared block"""
    n = len(models)
    lo, hi, per_rank = local_slice(n, hg.world, hg.rank)
    t_block, np_block = hg.block()
    if hi > lo:
        t_block[lo:hi].copy_(local_compute(models[lo:hi]), non_blocking=True)
    if stream_sync is not None:
        stream_sync()
    hg.publish_and_wait()
    return np_block[:n]


def sharded_evaluate(local_compute, models, n_nu, group=None, device="cpu", gather="all",
                     buffers=None):
    """Run `local_compute(models[lo:hi]) -> tensor [hi-lo, n_nu]` on this rank's shard and join the
    blocks with ONE collective.  Without an initialised process group this is a plain call.
    gather = "all":   all-gather, every rank returns the full [n, n_nu] (default);
             "rank0": the same collective, but only rank 0 returns the result (the others None) -- the
                      caller on the other ranks skips the device->host copy;
             "none":  no collective: every rank returns its own shard [hi-lo, n_nu].
    (gather="host" -- the batch assembled in node-shared host memory, no device collective -- is a mode
    of the host-level calls, see HostGather.)"""
    import torch.distributed as dist
    if gather not in GATHER_MODES:
        raise ValueError(f"gather must be one of {GATHER_MODES}")
    if gather == "host":
        raise ValueError("gather='host' assembles the batch in host memory: use __call__ / "
                         "host_sharded_evaluate, not the device-side evaluate")
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if world == 1:
        return local_compute(models)
    rank = dist.get_rank(group)
    n = len(models)
    lo, hi, per_rank = local_slice(n, world, rank)
    if gather == "none":
        return local_compute(models[lo:hi]) if hi > lo else None
    buffers = GatherBuffers() if buffers is None else buffers
    block, out = buffers.get(per_rank, n_nu, world, device, True)
    if hi > lo:
        block[: hi - lo] = local_compute(models[lo:hi])
    dist.all_gather_into_tensor(out, block, group=group)    # rows >= n are padding, cut below
    if gather == "rank0" and rank != 0:
        return None
    return out[:n]


def nu_indices(n_nu, world_size, rank):
    """indices of the frequencies owned by `rank` and the padded per-rank count (config 4: the nu
    grid is sharded, every other grid is replicated).  Round-robin, not contiguous: the cost of a
    frequency falls steeply along the grid (fewer and fewer gamma points stay unmasked towards the
    kinematic limit), so contiguous slices would leave the last ranks idle."""
    per_rank = -(-n_nu // world_size)
    return np.arange(rank, n_nu, world_size), per_rank


def nu_sharded_evaluate(local_compute, nu, n_models, group=None, device="cpu", buffers=None):
    """Frequency sharding: `local_compute(nu[idx]) -> tensor [n_models, len(idx)]` on this rank's
    frequencies (idx = rank, rank + world, ...), then ONE all-gather; returns [n_models, n_nu] on
    every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if world == 1:
        return local_compute(nu)
    rank = dist.get_rank(group)
    n = len(nu)
    idx, per_rank = nu_indices(n, world, rank)
    buffers = GatherBuffers() if buffers is None else buffers
    block, out = buffers.get(per_rank, n_models, world, device, True)             # [nu, model]
    if idx.size:
        block[: idx.size] = local_compute(nu[idx]).transpose(0, 1)
    dist.all_gather_into_tensor(out, block, group=group)
    # out[r * per_rank + j] is frequency j * world + r: undo the interleave
    out = out.view(world, per_rank, n_models).transpose(0, 1).reshape(world * per_rank, n_models)
    return out[:n].transpose(0, 1).contiguous()


# ---- the GPU plan -----------------------------------------------------------------------------
class SedBatch:
    """One radiative process on fixed grids, evaluated for batches of models on one GPU
    (or sharded over the ranks of `group`)."""

    def __init__(self, process, nu, ne_kind="BrokenPowerLaw", *, gamma=None, mu=None, phi=None,
                 nu_seed=None, mu_size=100, l_size=50, ssa=False, integrator="trapz",
                 device=None, group=None, shard="models"):
        import torch
        from . import grids
        if process not in PROCESSES:
            raise ValueError(f"process must be one of {PROCESSES}")
        if not torch.cuda.is_available():
            raise _lib.AgnpyB200Error("SedBatch needs a CUDA device: agnpy_b200 has no CPU fallback")
        lib()
        self.torch = torch
        self.process = process
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.group = group
        if shard not in ("models", "nu"):
            raise ValueError("shard must be 'models' or 'nu'")
        self.shard = shard
        self.nu_full = np.ascontiguousarray(np.ravel(nu), dtype=np.float64)
        if shard == "nu":  # this rank keeps only its slice of the frequency grid on the device
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                idx, _ = nu_indices(self.nu_full.size, dist.get_world_size(group), dist.get_rank(group))
                nu = self.nu_full[idx] if idx.size else self.nu_full[:1]
                self._nu_empty = idx.size == 0
            else:
                self._nu_empty = False
        self.ne_kind = _NE_KINDS[ne_kind] if isinstance(ne_kind, str) else int(ne_kind)
        self.integ = grids.integrator_code(integrator)
        self.ssa = int(bool(ssa))
        self.mu_size = int(mu_size)
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.device)
        self.nu = dev(np.ravel(nu))
        self.n_nu = self.nu.numel()
        if gamma is None:
            gamma = grids.gamma_e_to_integrate
        self.gamma = dev(grids.as_grid(gamma, "gamma"))
        self.mu = dev(grids.mu_to_integrate if mu is None else grids.as_grid(mu, "mu"))
        self.phi = dev(grids.phi_to_integrate if phi is None else grids.as_grid(phi, "phi"))
        if process in _EC:  # mirror-symmetric phi grid: the EC kernels may fold it
            self.integ |= grids.phi_fold_flag(grids.phi_to_integrate if phi is None else phi)
        # the np.trapz SSC kernel locates each electron's seed band by a monotone search: the seed
        # grid must be strictly ascending (and positive), like every other grid
        nu_seed = grids.nu_to_integrate if nu_seed is None else grids.as_grid(nu_seed, "nu_seed")
        if not nu_seed[0] > 0:
            raise ValueError("nu_seed: frequencies must be positive")
        self.nu_seed = dev(nu_seed)
        # normalisation grid of the thermal disk SED (targets/targets.py:151, 393-415)
        self.nu_bb_norm = dev(grids.nu_to_integrate_targets)
        self.l_unit = dev(np.logspace(-5, 5, int(l_size)) if process in _TAU_OFF_AXIS
                          else np.logspace(0, 5, int(l_size)))
        self._pinned_in = self._pinned_out = self._d_models = self._d_out = None
        self._gather = GatherBuffers()
        self._h2d_done = None       # event after the last pinned->device copy of the model table

    # -- buffers sized for the largest batch seen so far
    def _buffers(self, n):
        torch = self.torch
        if self._d_models is None or self._d_models.shape[0] < n:
            self._pinned_in = torch.empty(n, _lib.MODEL_DOUBLES, dtype=torch.float64).pin_memory()
            self._pinned_out = torch.empty(n, self.n_nu, dtype=torch.float64).pin_memory()
            self._d_models = torch.empty(n, _lib.MODEL_DOUBLES, dtype=torch.float64, device=self.device)
            self._d_out = torch.empty(n, self.n_nu, dtype=torch.float64, device=self.device)
        return self._pinned_in[:n], self._pinned_out[:n], self._d_models[:n], self._d_out[:n]

    def launch(self, d_models, d_out, stream=None):
        """enqueue the kernels for device-resident models [n, 22] -> d_out [n, n_nu]"""
        torch = self.torch
        n = d_models.shape[0]
        st = (torch.cuda.current_stream(self.device) if stream is None else stream).cuda_stream
        L, p = lib(), self.process
        if p == "synch":
            rc = L.agnpy_b200_synch_sed(n, d_models.data_ptr(), self.ne_kind, self.nu.data_ptr(),
                                        self.n_nu, self.gamma.data_ptr(), self.gamma.numel(), None,
                                        None, self.ssa, self.integ, d_out.data_ptr(), None, st)
        elif p == "ssc":
            rc = L.agnpy_b200_ssc_sed(n, d_models.data_ptr(), self.ne_kind, self.nu.data_ptr(),
                                      self.n_nu, self.nu_seed.data_ptr(), self.nu_seed.numel(),
                                      self.gamma.data_ptr(), self.gamma.numel(), None, None,
                                      self.ssa, self.integ, d_out.data_ptr(), st)
        elif p in _EC:
            v = _EC[p]
            mu_ptr = self.mu.data_ptr() if v in (_lib.EC_ISO_MONO, _lib.EC_BLR) else None
            n_mu = self.mu_size if v == _lib.EC_SS_DISK else self.mu.numel()
            rc = L.agnpy_b200_ec_sed(v, n, d_models.data_ptr(), self.ne_kind, self.nu.data_ptr(),
                                     self.n_nu, self.gamma.data_ptr(), self.gamma.numel(), mu_ptr,
                                     n_mu, self.phi.data_ptr(), self.phi.numel(), None, self.integ,
                                     d_out.data_ptr(), st)
        elif p in _BB:
            rc = L.agnpy_b200_bb_sed(_BB[p], n, d_models.data_ptr(), self.nu.data_ptr(), self.n_nu,
                                     self.nu_bb_norm.data_ptr(), self.nu_bb_norm.numel(), 100,
                                     d_out.data_ptr(), st)
        else:
            v = _TAU[p]
            mu_ptr = self.mu.data_ptr() if v in (_lib.TAU_BLR, _lib.TAU_BLR_MU_S) else None
            n_a = self.mu_size if v == _lib.TAU_SS_DISK else self.mu.numel()
            rc = L.agnpy_b200_tau(v, n, d_models.data_ptr(), self.nu.data_ptr(), self.n_nu, mu_ptr,
                                  n_a, self.phi.data_ptr(), self.phi.numel(),
                                  self.l_unit.data_ptr(), self.l_unit.numel(), d_out.data_ptr(), st)
        check(rc, f"agnpy_b200 {p}")
        return d_out

    def _local(self, models):
        """host models -> device result for this rank's shard (H2D from pinned memory).
        The returned tensor is the plan's own output buffer: valid until the next call on this plan."""
        n = len(models)
        pin_in, _, d_models, d_out = self._buffers(n)
        if self._h2d_done is not None:
            # the previous call's pinned->device copy may still be queued behind its kernels:
            # never rewrite the staging block under it
            self._h2d_done.synchronize()
        pin_in.numpy()[...] = np.ascontiguousarray(models).view(np.float64).reshape(n, -1)
        d_models.copy_(pin_in, non_blocking=True)
        if self._h2d_done is None:
            self._h2d_done = self.torch.cuda.Event()
        self._h2d_done.record(self.torch.cuda.current_stream(self.device))
        return self.launch(d_models, d_out)

    def evaluate_device(self, models, gather="all", clone=False):
        """as __call__ but returns the (gathered) device tensor without the D2H copy.  The tensor is
        a buffer of the plan, overwritten by the next call; pass clone=True to own the result."""
        if self.shard == "nu":
            def local(nu_part):  # the plan already holds this rank's slice
                return self._local(models)[:, : len(nu_part)]
            out = nu_sharded_evaluate(local, self.nu_full, len(models), self.group, self.device,
                                      self._gather)
        else:
            out = sharded_evaluate(self._local, models, self.n_nu, self.group, self.device, gather,
                                   self._gather)
        return out.clone() if (clone and out is not None) else out

    def _to_host(self, out, copy):
        """device [n, w] -> pinned host memory; returns a view of the plan's pinned block (valid until
        the next call) or, with copy=True, an array the caller owns"""
        torch = self.torch
        n, w = out.shape
        if (self._pinned_out is None or self._pinned_out.shape[0] < n
                or self._pinned_out.shape[1] != w):
            self._pinned_out = torch.empty(n, w, dtype=torch.float64).pin_memory()
        host = self._pinned_out[:n]
        host.copy_(out, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        host = host.numpy()
        return host.copy() if copy else host

    def __call__(self, models, gather="all", copy=True):
        """models (host) -> numpy [n_models, n_nu].
        gather: see `sharded_evaluate` ("rank0": ranks != 0 return None and copy nothing back;
                "none": every rank gets only its own shard) and `HostGather` ("host": every rank
                copies only its own shard device->host, into a block of host memory shared by the
                ranks of the node; everybody returns the whole batch, no device collective).
        copy=False returns a view of the plan's pinned result block, valid until the next call on
        this plan (saves one host memcpy of the whole result per call)."""
        if models.dtype != _lib.MODEL_DTYPE:
            raise TypeError("models must be a numpy array of agnpy_b200._lib.MODEL_DTYPE")
        with self.torch.cuda.device(self.device):
            if gather == "host" and self.shard == "models" and _world(self.group) > 1:
                return _host_gather_call(self, self._local, models, self.n_nu, copy)
            out = self.evaluate_device(models, "all" if gather == "host" else gather)
            if out is None:
                self.torch.cuda.current_stream(self.device).synchronize()
                return None
            return self._to_host(out, copy)


def _world(group):
    import torch.distributed as dist
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def _host_gather_call(owner, local_compute, items, n_nu, copy):
    """shared body of SedBatch / FitSedBatch __call__(gather="host"); `owner` keeps the HostGather"""
    import torch.distributed as dist
    world = _world(owner.group)
    per_rank = -(-len(items) // world)
    hg = getattr(owner, "_host_gather", None)
    if hg is None or hg.rows < per_rank * world or hg.cols != n_nu:
        if hg is not None:
            dist.barrier(group=owner.group)
            hg.close()
        hg = owner._host_gather = HostGather(owner.group, per_rank * world, n_nu)
    sync = owner.torch.cuda.current_stream(owner.device).synchronize
    out = host_sharded_evaluate(local_compute, items, n_nu, hg, sync)
    return out.copy() if copy else out


class GraphedSum:
    """CUDA graph of `sum_k plans[k](models_k)` (+ optional EBL table) for a FIXED batch size on one
    GPU: the pinned->device copy of the model tables, every kernel of every component, the
    additions and the device->pinned copy of the result are captured once and replayed per call.
    A fit / MCMC step with a few dozen parameter vectors is launch-bound (about 20 launches of a few
    microseconds each); the replay submits them as one unit.

    The graph is captured on a private stream: the library's scratch buffer is keyed by stream, so
    the pointers baked into the graph are never reallocated by other calls."""

    def __init__(self, plans, n_models, ebl_tables=None):
        first = next(iter(plans.values()))
        torch = self.torch = first.torch
        self.device, self.n, self.n_nu = first.device, int(n_models), first.n_nu
        self.plans, self.names, self.ebl = plans, list(plans), ebl_tables
        k = len(self.names)
        f64 = torch.float64
        self.pin = torch.empty(k, self.n, _lib.MODEL_DOUBLES, dtype=f64).pin_memory()
        self.pin_z = torch.empty(self.n, dtype=f64).pin_memory()
        self.pin_out = torch.empty(self.n, self.n_nu, dtype=f64).pin_memory()
        self.d_models = torch.empty(k, self.n, _lib.MODEL_DOUBLES, dtype=f64, device=self.device)
        self.d_z = torch.empty(self.n, dtype=f64, device=self.device)
        self.d_part = torch.empty(self.n, self.n_nu, dtype=f64, device=self.device)
        self.d_total = torch.empty(self.n, self.n_nu, dtype=f64, device=self.device)
        self.pin.zero_()
        self.pin_z.zero_()
        self.stream = torch.cuda.Stream(self.device)
        with torch.cuda.device(self.device):
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.stream):
                for _ in range(2):            # warm-up: scratch allocation, function attributes
                    self._enqueue()
            self.stream.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=self.stream):
                self._enqueue()

    def _enqueue(self):
        torch = self.torch
        self.d_models.copy_(self.pin, non_blocking=True)
        for k, name in enumerate(self.names):
            self.plans[name].launch(self.d_models[k], self.d_total if k == 0 else self.d_part)
            if k:
                self.d_total.add_(self.d_part)
        if self.ebl is not None:
            self.d_z.copy_(self.pin_z, non_blocking=True)
            t, nu_dev = self.ebl, self.plans[self.names[0]].nu
            st = torch.cuda.current_stream(self.device).cuda_stream
            check(lib().agnpy_b200_ebl_absorption(
                self.n, self.d_z.data_ptr(), nu_dev.data_ptr(), self.n_nu, t["e"].data_ptr(),
                t["e"].numel(), t["z"].data_ptr(), t["z"].numel(), t["v"].data_ptr(), 1,
                self.d_total.data_ptr(), st), "agnpy_b200_ebl_absorption")
        self.pin_out.copy_(self.d_total, non_blocking=True)

    def __call__(self, models_by_name):
        """models_by_name[name]: numpy array of _lib.MODEL_DTYPE with exactly n_models rows"""
        for k, name in enumerate(self.names):
            m = models_by_name[name]
            if len(m) != self.n or m.dtype != _lib.MODEL_DTYPE:
                raise ValueError(f"graph captured for {self.n} models of MODEL_DTYPE per component")
            self.pin[k].numpy()[...] = np.ascontiguousarray(m).view(np.float64).reshape(self.n, -1)
        if self.ebl is not None:
            self.pin_z.numpy()[...] = models_by_name[self.names[0]]["z"]
        torch = self.torch
        with torch.cuda.device(self.device):
            self.graph.replay()
            torch.cuda.current_stream(self.device).synchronize()
        return self.pin_out.numpy().copy()


def evaluate_sed_batch(process, nu, models, ne_kind="BrokenPowerLaw", **kwargs):
    """one-shot convenience: build a plan and evaluate `models` (see SedBatch)"""
    return SedBatch(process, nu, ne_kind, **kwargs)(models)
