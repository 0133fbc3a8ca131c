"""Model-sharded multi-GPU plumbing.

The hot path shards by server: rank r owns servers ``s % world == r`` (or, in the weak-scaling
bench, its own 1k models); the small accelerator / capacity tables are replicated and the sizing has
**no data-path collective**.  What crosses ranks:

* unlimited capacity (SolveUnlimited is per server): one all-reduce(sum) of per-shard partials — the by-type GPU
  counts and costs of System.AllocateByType (pkg/core/system.go:271-299) plus a few counters;
* limited capacity (SolveGreedy, pkg/solver/greedy.go:35-105, orders ALL servers against shared per-type pools): one
  all-gather of the candidate allocations (37 B per (server, accelerator) pair), after which every rank runs the same
  sweep on the merged set (wva_set_candidates + wva_solve) and holds the same global solution — the sizing, >99 % of
  the time, stays sharded.

Both go over NCCL (NVLink/NVSwitch) through torch.distributed; on CPU the same code runs over gloo in the tests.
torch is plumbing here (process group + device buffer), not compute.
"""
from __future__ import annotations

import numpy as np


def solution_partials(sol: dict, frontier=None) -> np.ndarray:
    """[type_count (T) | type_cost (T) | n_allocated | n_unallocated | total_replicas | frontier_sum] as float64.

    Counts are integers far below 2**53, so the float64 sum across ranks is exact for them.
    """
    tc = np.asarray(sol["type_count"], dtype=np.float64)
    tk = np.asarray(sol["type_cost"], dtype=np.float64)
    state = np.asarray(sol["state"])
    extra = [float((state == 1).sum()), float((state == 0).sum()),
             float(np.asarray(sol["num_replicas"], dtype=np.int64)[state == 1].sum()),
             float(np.asarray(frontier, dtype=np.int64).sum()) if frontier is not None else 0.0]
    return np.concatenate([tc, tk, extra])


def split_partials(vec: np.ndarray, n_types: int) -> dict:
    vec = np.asarray(vec, dtype=np.float64)
    T = n_types
    return {"type_count": np.rint(vec[:T]).astype(np.int64), "type_cost": vec[T:2 * T],
            "n_allocated": int(round(vec[2 * T])), "n_unallocated": int(round(vec[2 * T + 1])),
            "total_replicas": int(round(vec[2 * T + 2])), "frontier_sum": int(round(vec[2 * T + 3]))}


def all_reduce_partials(vec: np.ndarray, device=None) -> np.ndarray:
    """Sum the per-shard partial vector over all ranks (identity when no process group is up)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return np.asarray(vec, dtype=np.float64)
    t = torch.from_numpy(np.ascontiguousarray(vec, dtype=np.float64))
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def shard_indices(n_servers: int, rank: int, world: int) -> np.ndarray:
    """servers owned by `rank` under the canonical round-robin partition"""
    return np.arange(rank, n_servers, world)


# ---- limited-capacity solve over model shards ---------------------------------------------------------------------------
_CAND_FIELDS = (("state", np.uint8), ("num_replicas", np.int32), ("batch_size", np.int32), ("cost", np.float32),
                ("value", np.float32), ("itl", np.float32), ("ttft", np.float32), ("rho", np.float32),
                ("max_arrv_rate", np.float32), ("n_solves", np.int32))


def pack_candidates(cand: dict, rows: int, n_acc: int) -> np.ndarray:
    """field-major byte image of a shard's candidate arrays, padded to `rows` servers"""
    parts = []
    for k, dt in _CAND_FIELDS:
        a = np.zeros((rows, n_acc), dtype=dt)
        src = np.asarray(cand[k], dtype=dt).reshape(-1, n_acc)
        a[: src.shape[0]] = src
        parts.append(a.reshape(-1).view(np.uint8))
    return np.concatenate(parts) if parts else np.zeros(0, np.uint8)


def unpack_candidates(buf: np.ndarray, rows: int, n_acc: int) -> dict:
    out, off = {}, 0
    for k, dt in _CAND_FIELDS:
        nb = rows * n_acc * np.dtype(dt).itemsize
        out[k] = np.ascontiguousarray(buf[off:off + nb]).view(dt).reshape(rows, n_acc)
        off += nb
    return out


def gather_candidates(cand: dict, n_servers: int, n_acc: int, rank: int, world: int, device=None) -> dict:
    """All-gather the shards' candidate arrays into [n_servers, n_acc] arrays in the server order of the full system
    (shard r holds servers r, r + world, ...: server j * world + r is row j of shard r, so the merge is one transpose
    per field).  One collective of ceil(S / world) * A * 37 bytes per rank."""
    import torch
    import torch.distributed as dist
    rows = (n_servers + world - 1) // world
    mine = pack_candidates(cand, rows, n_acc)
    if world == 1 or not (dist.is_available() and dist.is_initialized()):
        allb = mine.reshape(1, -1)
    else:
        t = torch.from_numpy(mine)
        if device is not None:
            t = t.to(device)
        out = torch.empty(world * t.numel(), dtype=torch.uint8, device=t.device)     # flat: gloo accepts no other shape
        dist.all_gather_into_tensor(out, t)
        allb = out.cpu().numpy().reshape(world, -1)
    full, off = {}, 0
    for k, dt in _CAND_FIELDS:
        nb = rows * n_acc * np.dtype(dt).itemsize
        g = np.ascontiguousarray(allb[:, off:off + nb]).view(dt).reshape(world, rows, n_acc)      # [rank][row][acc]
        full[k] = np.ascontiguousarray(g.transpose(1, 0, 2)).reshape(rows * world, n_acc)[:n_servers]
        off += nb
    return full


class ShardError(RuntimeError):
    """Raised on EVERY rank when some rank could not size its shard."""


def _agree(err, rank: int, world: int, device=None):
    """All-reduce (MAX) a failure flag; raise on every rank when any rank failed."""
    bad = 1 if err is not None else 0
    if world > 1:
        import torch
        import torch.distributed as dist
        flag = torch.tensor([bad, rank if bad else -1], dtype=torch.int32, device=device if device is not None else "cpu")
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        bad, who = int(flag[0].item()), int(flag[1].item())
    else:
        who = rank
    if bad:
        if err is not None:
            raise ShardError(f"rank {rank}: {err}") from err
        raise ShardError(f"rank {who} failed to size its shard (this rank, {rank}, was fine)")


def solve_sharded(engine, sysd: dict, rank: int, world: int, device=None, shard_fn=None, timings: dict | None = None):
    """Manager.Optimize (pkg/manager/manager.go:21-27) over `world` ranks for either capacity mode: size the rank's
    shard, all-gather the candidates, run the allocator on the merged set.  Every rank returns the global solution.
    `engine` is an Engine (or, in the CPU tests, a stand-in with the same six methods); `timings` collects host
    wall-clock milliseconds per step."""
    import time
    if shard_fn is None:
        from .synth import shard_system as shard_fn
    S, A = int(sysd["n_servers"]), int(sysd["n_acc"])
    t = [time.perf_counter()]

    def lap(name):
        t.append(time.perf_counter())
        if timings is not None:
            timings[name] = (t[-1] - t[-2]) * 1e3

    shard, idx = shard_fn(sysd, rank, world)
    assert np.array_equal(idx, shard_indices(S, rank, world))
    lap("shard")
    # A rank whose shard fails to load or to size (a limit, out of memory) must not leave the others blocked in the
    # all-gather: every rank learns the worst status first and all of them raise.
    err, cand = None, None
    try:
        engine.load_system(shard)
        engine.calculate()
        lap("load+calculate")
        cand = engine.candidates()
        lap("candidates_d2h")
    except Exception as e:                      # noqa: BLE001 — re-raised below, on every rank
        err = e
    _agree(err, rank, world, device)
    full = gather_candidates(cand, S, A, rank, world, device)
    lap("all_gather+merge")
    engine.load_system(sysd)
    engine.set_candidates(full)
    lap("load+set_candidates")
    engine.solve()
    lap("solve")
    sol = engine.solution()
    lap("solution_d2h")
    return sol
