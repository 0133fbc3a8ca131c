"""World-size-2 gloo test of the multi-GPU plumbing (no GPU): each rank sizes + solves its model shard
with the oracle standing in for the device (the test checks the SHARDING and the all-reduce, not the
kernels), all-reduces the per-shard partials, and both ranks must end with the global by-type totals."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "llm-d-workload-variant-autoscaler_b200"


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    pkg_synth = importlib.import_module(PKG + ".synth")
    sharding = importlib.import_module(PKG + ".sharding")
    from tests import oracle_lib
    orc = oracle_lib.load()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        d = pkg_synth.queue_system(24, 4, 16, stream=81)
        sd, idx = pkg_synth.shard_system(d, rank, world)
        assert np.array_equal(idx, sharding.shard_indices(24, rank, world))
        cand = orc.calculate(sd, nthreads=1)
        sol = orc.solve(sd, cand)
        vec = sharding.solution_partials(sol, frontier=np.zeros((len(idx), 4), np.int32))
        tot = sharding.all_reduce_partials(vec)
        q.put((rank, tot))
    finally:
        dist.destroy_process_group()


def test_sharded_partials_all_reduce_gloo(pkg, oracle):
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    d = pkg.synth.queue_system(24, 4, 16, stream=81)
    full = oracle.solve(d, oracle.calculate(d))
    ref = pkg.sharding.split_partials(pkg.sharding.solution_partials(full), d["n_types"])
    for r in range(world):
        tot = pkg.sharding.split_partials(got[r], d["n_types"])
        assert np.array_equal(tot["type_count"], ref["type_count"])
        np.testing.assert_allclose(tot["type_cost"], ref["type_cost"], rtol=1e-12)
        assert tot["n_allocated"] == ref["n_allocated"] and tot["total_replicas"] == ref["total_replicas"]


def test_partials_roundtrip(pkg):
    sol = {"type_count": np.array([3, 0, 7]), "type_cost": np.array([1.5, 0.0, 2.25]), "state": np.array([1, 0, 1, 2]),
           "num_replicas": np.array([2, 0, 5, 0])}
    p = pkg.sharding.split_partials(pkg.sharding.solution_partials(sol, frontier=np.array([[1, 2], [3, 0]])), 3)
    assert p["type_count"].tolist() == [3, 0, 7] and p["n_allocated"] == 2 and p["n_unallocated"] == 1
    assert p["total_replicas"] == 7 and p["frontier_sum"] == 6
