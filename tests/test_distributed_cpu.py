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


class _OracleEngine:
    """the five Engine methods solve_sharded uses, with the oracle standing in for the device"""

    def __init__(self, orc):
        self.orc = orc

    def load_system(self, d):
        self.d, self.cand = d, None

    def calculate(self):
        self.cand = self.orc.calculate(self.d, nthreads=1)

    def candidates(self):
        return self.cand

    def set_candidates(self, cand):
        self.cand = cand

    def solve(self):
        self.sol = self.orc.solve(self.d, self.cand)

    def solution(self):
        return self.sol


def _limited_system(pkg_synth, orc, policy, delayed):
    d = pkg_synth.queue_system(23, 5, 16, stream=83, saturation_policy=policy, delayed_best_effort=delayed)
    un = orc.solve(d, orc.calculate(d, nthreads=1))
    return pkg_synth.limit_capacity(d, un["type_count"], 0.5)     # half of the unconstrained demand: the pools bind


def _greedy_worker(rank, world, port, q):
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
        out = []
        for policy, delayed in ((0, False), (1, False), (2, True), (3, False)):
            d = _limited_system(pkg_synth, orc, policy, delayed)
            sol = sharding.solve_sharded(_OracleEngine(orc), d, rank, world)
            out.append({k: np.asarray(v) for k, v in sol.items()})
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_sharded_limited_greedy_all_gather_gloo(pkg, oracle):
    """Limited capacity: the greedy sweep needs every server, so the shards' candidates are all-gathered and each rank
    solves the merged set — 23 servers over 2 ranks (ragged shards), all four saturation policies."""
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_greedy_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for i, (policy, delayed) in enumerate(((0, False), (1, False), (2, True), (3, False))):
        d = _limited_system(pkg.synth, oracle, policy, delayed)
        ref = oracle.solve(d, oracle.calculate(d))
        if policy == 0:
            assert (np.asarray(ref["state"]) == 1).any() and (np.asarray(ref["state"]) == 0).any()   # the cap binds
        for r in range(world):
            for k, v in ref.items():
                a, b = np.asarray(got[r][i][k]), np.asarray(v)
                assert np.array_equal(a.view(np.uint8) if a.dtype.kind == "f" else a, b.view(np.uint8) if b.dtype.kind == "f" else b), (policy, r, k)


class _FailingEngine(_OracleEngine):
    def calculate(self):
        raise RuntimeError("WVA_ERR_LIMIT: one pair over the limit")


def _failing_worker(rank, world, port, q):
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
        d = _limited_system(pkg_synth, orc, 0, False)
        eng = _FailingEngine(orc) if rank == 1 else _OracleEngine(orc)
        try:
            sharding.solve_sharded(eng, d, rank, world)
            q.put((rank, "no error"))
        except sharding.ShardError as e:
            q.put((rank, str(e)))
    finally:
        dist.destroy_process_group()


def test_sharded_solve_raises_on_every_rank_when_one_fails():
    """a rank whose calculate() raises must not leave the others blocked in the all-gather."""
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_failing_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert "rank 1" in got[0] and "this rank, 0, was fine" in got[0]
    assert "WVA_ERR_LIMIT" in got[1]


def test_candidate_pack_roundtrip(pkg):
    g = np.random.default_rng(5)
    sh = pkg.sharding
    cand = {k: (g.integers(0, 3, (7, 3)).astype(dt) if np.dtype(dt).kind in "ui" else g.random((7, 3)).astype(dt)) for k, dt in sh._CAND_FIELDS}
    back = sh.unpack_candidates(sh.pack_candidates(cand, 9, 3), 9, 3)
    for k, _ in sh._CAND_FIELDS:
        assert np.array_equal(back[k][:7], cand[k]) and not back[k][7:].any()
    one = sh.gather_candidates(cand, 7, 3, 0, 1)
    for k, _ in sh._CAND_FIELDS:
        assert np.array_equal(one[k], cand[k])


def test_partials_roundtrip(pkg):
    sol = {"type_count": np.array([3, 0, 7]), "type_cost": np.array([1.5, 0.0, 2.25]), "state": np.array([1, 0, 1, 2]),
           "num_replicas": np.array([2, 0, 5, 0])}
    p = pkg.sharding.split_partials(pkg.sharding.solution_partials(sol, frontier=np.array([[1, 2], [3, 0]])), 3)
    assert p["type_count"].tolist() == [3, 0, 7] and p["n_allocated"] == 2 and p["n_unallocated"] == 1
    assert p["total_replicas"] == 7 and p["frontier_sum"] == 6
