"""Model-sharded multi-GPU plumbing.

The hot path shards by server: rank r owns servers ``s % world == r`` (or, in the weak-scaling
bench, its own 1k models); the small accelerator / capacity tables are replicated and **no data-path
collective exists**.  What crosses ranks is one all-reduce(sum) of per-shard partials — the by-type
GPU counts and costs of System.AllocateByType (pkg/core/system.go:271-299) plus a few counters — over
NCCL (NVLink/NVSwitch) through torch.distributed; on CPU the same code runs over gloo in the tests.
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
