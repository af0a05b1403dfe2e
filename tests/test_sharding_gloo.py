"""The N>1 path on CPU: batch sharding + one all-gather over a world_size-2 gloo group.
The GPU compute is replaced by the oracle (tests may use it), so what is exercised is the
host logic of agnpy_b200.batch: partition, padding, gather order, trimming."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from agnpy_b200 import _core, _lib
from agnpy_b200.batch import local_slice, nu_sharded_evaluate, partition, sharded_evaluate


def test_partition_arithmetic():
    assert partition(10, 4) == (3, 12)
    assert partition(8, 8) == (1, 8)
    assert partition(1, 8) == (1, 8)
    covered = []
    for r in range(4):
        lo, hi, per = local_slice(10, 4, r)
        assert per == 3
        covered += list(range(lo, hi))
    assert covered == list(range(10))
    assert local_slice(1, 8, 5) == (1, 1, 1)          # empty shard


def _models(n):
    models = _lib.make_models(n)
    for i in range(n):
        _core.model_row(models, i, z=0.05 * i, d_L=1e27, delta_D=10 + i, B=0.5, R_b=1e16,
                        ne=[1e-8, 2.0 + 0.05 * i, 10.0, 1e5, 0, 0, 0, 0])
    return models


def _cpu_compute(nu):
    from oracle import agnpy_oracle as orc

    def f(models):
        out = [orc.synch_sed_flux(nu, m["z"], m["d_L"], m["delta_D"], m["B"], m["R_b"], "PowerLaw",
                                  tuple(m["ne"][:4]), gamma=np.logspace(1, 5, 60)) for m in models]
        return torch.from_numpy(np.array(out).reshape(len(models), nu.size))
    return f


def _worker(rank, world, port, n_models, q, gather="all"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nu = np.logspace(9, 18, 12)
    from agnpy_b200.batch import GatherBuffers
    buffers = GatherBuffers()
    for _ in range(2):      # the second call reuses the staging buffers of the first
        out = sharded_evaluate(_cpu_compute(nu), _models(n_models), nu.size, device="cpu",
                               gather=gather, buffers=buffers)
    q.put((rank, None if out is None else out.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_models", [5, 4, 1])
def test_sharded_evaluate_world2_gloo(n_models):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_models, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    nu = np.logspace(9, 18, 12)
    full = _cpu_compute(nu)(_models(n_models)).numpy()
    for r in range(2):
        assert results[r].shape == (n_models, nu.size)
        np.testing.assert_array_equal(results[r], full)   # every rank holds the whole batch


@pytest.mark.parametrize("gather", ["rank0", "none"])
def test_gather_modes_world2_gloo(gather):
    """gather="rank0": the same collective, only rank 0 returns the batch (the others skip their
    device->host copy); gather="none": no collective, every rank returns its own shard"""
    n_models = 5
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_models, q, gather)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    nu = np.logspace(9, 18, 12)
    full = _cpu_compute(nu)(_models(n_models)).numpy()
    if gather == "rank0":
        np.testing.assert_array_equal(results[0], full)
        assert results[1] is None
    else:
        np.testing.assert_array_equal(results[0], full[:3])     # ceil(5/2) models on rank 0
        np.testing.assert_array_equal(results[1], full[3:])
    with pytest.raises(ValueError):
        sharded_evaluate(lambda m: None, _models(2), 12, gather="some")


def _host_worker(rank, world, port, n_models, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from agnpy_b200.batch import HostGather, host_sharded_evaluate
    nu = np.logspace(9, 18, 12)
    hg = HostGather(None, 2 * (-(-n_models // world)), nu.size, register=False)
    outs = []
    for scale in (1.0, 2.0, 3.0):       # three calls: both blocks of the ping-pong are reused
        f = _cpu_compute(nu)
        out = host_sharded_evaluate(lambda m: scale * f(m), _models(n_models), nu.size, hg)
        outs.append(out.copy())
    q.put((rank, outs))
    dist.barrier()
    hg.close()
    dist.destroy_process_group()


def test_host_gather_world2_gloo():
    """gather="host": every rank copies its own shard into node-shared host memory; after the
    sequence barrier every rank sees the whole batch (no collective on the data path)"""
    n_models = 5
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_host_worker, args=(r, 2, port, n_models, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    nu = np.logspace(9, 18, 12)
    full = _cpu_compute(nu)(_models(n_models)).numpy()
    for r in range(2):
        for k, scale in enumerate((1.0, 2.0, 3.0)):
            np.testing.assert_array_equal(results[r][k], scale * full)


def _nu_worker(rank, world, port, n_nu, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nu = np.logspace(9, 18, n_nu)
    models = _models(3)
    out = nu_sharded_evaluate(lambda part: _cpu_compute(part)(models), nu, len(models), device="cpu")
    q.put((rank, out.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_nu", [9, 8, 1])
def test_nu_sharded_evaluate_world2_gloo(n_nu):
    """BASELINE config 4 style frequency sharding: slices, padding, gather order"""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nu_worker, args=(r, 2, port, n_nu, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    nu = np.logspace(9, 18, n_nu)
    full = _cpu_compute(nu)(_models(3)).numpy()
    for r in range(2):
        assert results[r].shape == (3, n_nu)
        np.testing.assert_array_equal(results[r], full)
