"""Multi-GPU path THROUGH THE C-ABI (wva_comm_* / wva_group_*, NCCL inside the library).  Needs >= 2 GPUs: skipped on a
one-GPU box (the world-size-2 host logic runs on CPU over gloo in tests/test_distributed_cpu.py)."""
import importlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


needs2 = pytest.mark.skipif("_n_gpus() < 2", reason="needs two GPUs")

SOL_INT = ("state", "acc", "num_replicas", "batch_size")
F32 = ("cost", "value", "itl", "ttft", "rho", "max_arrv_rate")


def _same(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))


@needs2
@pytest.mark.parametrize("S", [401, 2000])
def test_group_optimize_equals_one_gpu(pkg, engine, oracle, S):
    """wva_group_optimize over 2 devices (one process, NCCL in the library) == the one-GPU path == the oracle, for the
    unlimited allocator (all-gather of the solution + all-reduce of the by-type partials) and the limited one (in-place
    all-gather of the candidate arena, greedy on every rank), ragged last block included."""
    d = pkg.synth.queue_system(S, 6, 32, stream=97)
    one = engine.optimize(d)
    cand = engine.candidates()
    with pkg.Group([0, 1]) as g:
        two = g.optimize(d)
        for k in SOL_INT + F32:
            assert _same(one[k], two[k]), k
        assert np.array_equal(one["type_count"], two["type_count"])
        np.testing.assert_allclose(one["type_cost"], two["type_cost"], rtol=1e-12)
        for pol, delayed in (("None", False), ("PriorityRoundRobin", True)):
            lim = pkg.synth.limit_capacity(d, one["type_count"], 0.55)
            lim["saturation_policy"] = pol; lim["delayed_best_effort"] = delayed
            a = engine.optimize(lim)
            b = g.optimize(lim)
            o = oracle.solve(lim, cand)
            for k in SOL_INT + F32:
                assert _same(a[k], b[k]) and _same(b[k], o[k]), (k, pol)
            assert np.array_equal(a["type_count"], b["type_count"]) and (a["state"] == 0).any()
            assert g.timing(0)["exchange_ms"] > 0


@needs2
def test_group_saturation_equals_one_gpu(pkg, engine, oracle):
    d = pkg.synth.saturation_batch(3001, 32, stream=4)
    one = engine.saturation_v1(d)
    with pkg.Group([0, 1]) as g:
        two = g.saturation_v1(d)
    for k in one:
        if k != "partials_all":
            assert _same(one[k], two[k]), k
    assert np.array_equal(two["partials_all"], one["partials"])


@needs2
def test_group_error_reaches_every_rank(pkg):
    """a rank whose sizing fails makes wva_solve fail on EVERY rank (status all-reduce) instead of hanging the others"""
    d = pkg.synth.queue_system(64, 4, 16, stream=5)
    d["perf_max_batch"] = d["perf_max_batch"].copy()
    d["perf_max_batch"][40:, :] = 70000            # second block only: N above the kernels' limit -> WVA_ERR_LIMIT there
    d["perf_at_tokens"] = d["srv_out_tokens"][:, None].repeat(4, axis=1).astype(np.int32)
    with pkg.Group([0, 1]) as g:
        with pytest.raises(pkg.WvaError):
            g.optimize(d)
        ok = pkg.synth.queue_system(64, 4, 16, stream=5)
        assert g.optimize(ok)["state"].shape == (64,)            # and the group is still usable


@needs2
def test_rank_level_comm_under_torchrun(pkg):
    """one process per GPU (how bench.py --gpus N runs): tools/run_multi_gpu.py compares the sharded result on every
    rank with a whole-system run, bit for bit"""
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29611",
                          os.path.join(ROOT, "tools", "run_multi_gpu.py"), "--servers", "3000", "--check"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    rep = json.loads(out.stdout.strip().splitlines()[-1])
    assert rep["identical_to_one_gpu"] and rep["world"] == 2
