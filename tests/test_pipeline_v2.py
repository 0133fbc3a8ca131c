"""V2 pipeline: token-capacity analyzer, cost-aware optimizer, enforcer.

CPU: the oracle (oracle/pipeline_v2.hpp) against the exact-value cases of the reference's own tests, and the host
mirrors (pipeline.py: k2 history, capacity store, record plumbing) driven through an engine stand-in whose three batch
calls are answered by the oracle — so the host logic is tested without a GPU.  GPU: the same cases and seeded random
batches through the C-ABI, device == oracle bit for bit (float64 compared as bits).

Reference tests mirrored: internal/engines/analyzers/saturation_v2/analyzer_test.go:31-663,
internal/engines/pipeline/cost_aware_optimizer_test.go:30-300, internal/engines/pipeline/enforcer_test.go:30-335.
"""
import numpy as np
import pytest

CFG = {"KvCacheThreshold": 0.8, "QueueLengthThreshold": 5, "KvSpareTrigger": 0.1, "QueueSpareTrigger": 3,
       "AnalyzerName": "saturation", "ScaleUpThreshold": 0.85, "ScaleDownBoundary": 0.70}


def rm(pod, variant, acc, cost, used, cap, q, ai, ao, hit=0.0):
    """makeReplicaMetrics (analyzer_test.go:697-723)"""
    return {"PodName": pod, "VariantName": variant, "AcceleratorName": acc, "Cost": cost, "TokensInUse": used,
            "TotalKvCapacityTokens": cap, "QueueLength": q, "AvgInputTokens": ai, "AvgOutputTokens": ao,
            "PrefixCacheHitRate": hit, "NumGpuBlocks": cap // 16, "BlockSize": 16, "ModelID": "test-model", "Namespace": "test-ns"}


def inp(metrics, states, queue=None):
    return {"ModelID": "test-model", "Namespace": "test-ns", "ReplicaMetrics": metrics, "VariantStates": states,
            "Config": dict(CFG), "SchedulerQueue": queue}


def st(name, cur, pending=0, gpus=1):
    return {"VariantName": name, "CurrentReplicas": cur, "PendingReplicas": pending, "GPUsPerReplica": gpus}


class OracleEngine:
    """Engine stand-in for the CPU tests: the three V2 batch calls answered by the oracle."""

    def __init__(self, oracle):
        self.o = oracle

    def saturation_v2(self, d):
        return self.o.saturation_v2(d)

    def cost_aware_optimize(self, d):
        return self.o.cost_aware_optimize(d)

    def enforce(self, d):
        return self.o.enforce(d)


# ---- reference cases, runnable on either engine --------------------------------------------------------------------------------
def _analyzer_cases(pkg, eng):
    A = lambda: pkg.pipeline.SaturationAnalyzerV2(eng)
    one = [st("variant-a", 1)]
    # k1/k2 interaction (analyzer_test.go:31-86)
    r = A().analyze(inp([rm("pod-1", "variant-a", "H100", 10.0, 5000, 16000, 0, 100, 50)], one))
    assert len(r["VariantCapacities"]) == 1 and r["VariantCapacities"][0]["PerReplicaCapacity"] == 12800.0
    assert A().analyze(inp([rm("pod-1", "variant-a", "H100", 10.0, 8000, 16000, 6, 100, 50)], one))["VariantCapacities"][0]["PerReplicaCapacity"] == 8000.0
    assert A().analyze(inp([rm("pod-1", "variant-a", "H100", 10.0, 4000, 16000, 10, 100, 50)], one))["VariantCapacities"][0]["PerReplicaCapacity"] == 4000.0
    # k2 history (:88-139)
    a = A()
    a.analyze(inp([rm("pod-1", "variant-a", "H100", 10.0, 8000, 16000, 6, 100, 50)], one))
    assert a.history["test-model|H100|short"] == [8000.0]
    r = a.analyze(inp([rm("pod-1", "variant-a", "H100", 10.0, 6000, 16000, 2, 100, 50)], one))
    assert r["VariantCapacities"][0]["PerReplicaCapacity"] == 8000.0
    # output-length bucketing (:141-170): a long-output replica does not see the short bucket's history
    r = a.analyze(inp([rm("pod-1", "variant-a", "H100", 10.0, 6000, 16000, 2, 100, 600)], one))
    assert r["VariantCapacities"][0]["PerReplicaCapacity"] == 12800.0
    # k2 derived from deployment params (:172-224): B=4096, S=256, I=500, O=100 -> 140800 > k1 -> k1 wins; small B binds
    a = A()
    a.store.update("test-ns", "test-model", "variant-a", {"AcceleratorName": "H100", "GpuCount": 1, "LearnedFrom": "deployment",
                                                          "EffectiveCapacity": 2048,
                                                          "VLLMParams": {"EffectiveMaxBatchedTokens": 2048, "MaxNumSeqs": 8}})
    r = a.analyze(inp([rm("pod-1", "variant-a", "H100", 10.0, 1000, 16000, 0, 500, 100)], one))
    # n_steady = min(2048*100/600, 8) = 8 -> k2 = 8 * 550 = 4400 < k1 = 12800
    assert r["VariantCapacities"][0]["PerReplicaCapacity"] == 4400.0
    assert a.store.get("test-ns", "test-model", "variant-a")["LearnedFrom"] == "live"
    assert a.store.get("test-ns", "test-model", "variant-a")["VLLMParams"]["MaxNumSeqs"] == 8       # params preserved
    # pending replicas (:226-263)
    r = A().analyze(inp([rm("pod-1", "variant-a", "H100", 10.0, 10000, 16000, 0, 100, 50)], [st("variant-a", 2, 1)]))
    assert r["VariantCapacities"][0]["ReplicaCount"] == 1 and r["RequiredCapacity"] >= 0
    # anticipated = 2 x 12800: required = 10000/0.85 - 25600 < 0 -> 0
    assert r["RequiredCapacity"] == 0.0
    r = A().analyze(inp([rm("pod-1", "variant-a", "H100", 10.0, 1000, 16000, 0, 100, 50)], [st("variant-a", 3, 1)]))
    assert r["TotalSupply"] == 2 * 12800.0 and r["SpareCapacity"] == 25600.0 - 1000 / 0.70
    # zero-replica variant with a live record (:266-286)
    a = A()
    a.store.update("test-ns", "test-model", "variant-b", {"AcceleratorName": "A100", "GpuCount": 1, "EffectiveCapacity": 9000, "LearnedFrom": "live"})
    r = a.analyze(inp([rm("pod-1", "variant-a", "H100", 10.0, 5000, 16000, 0, 100, 50)], [st("variant-a", 1), st("variant-b", 0)]))
    vb = [v for v in r["VariantCapacities"] if v["VariantName"] == "variant-b"][0]
    assert vb["PerReplicaCapacity"] == 9000.0 and vb["TotalCapacity"] == 0.0 and vb["ReplicaCount"] == 0
    # zero-replica variants (:266-442): live record, deployment-derived record + workload, bounds, fallback
    a = A()
    a.store.update("test-ns", "test-model", "variant-a", {"AcceleratorName": "H100", "GpuCount": 1, "EffectiveCapacity": 12000, "LearnedFrom": "live"})
    r = a.analyze(inp([], [st("variant-a", 0)]))
    assert len(r["VariantCapacities"]) == 1 and r["VariantCapacities"][0]["PerReplicaCapacity"] == 12000.0
    a = A()
    a.store.update("test-ns", "test-model", "variant-b", {"AcceleratorName": "A100", "GpuCount": 1, "EffectiveCapacity": 8192, "LearnedFrom": "deployment",
                                                          "VLLMParams": {"EffectiveMaxBatchedTokens": 8192, "MaxNumSeqs": 256}})
    r = a.analyze(inp([rm("pod-1", "variant-a", "H100", 10.0, 5000, 16000, 0, 500, 100)], [st("variant-a", 1), st("variant-b", 0)]))
    assert r["VariantCapacities"][1]["VariantName"] == "variant-b" and r["VariantCapacities"][1]["PerReplicaCapacity"] == 140800.0
    a = A()
    dp = {"GpuMemoryUtilization": 0.9, "BlockSize": 16, "KvCacheDtype": "auto", "TensorParallelSize": 1, "MaxNumSeqs": 256, "EffectiveMaxBatchedTokens": 8192}
    a.store.update("test-ns", "test-model", "variant-b", {"AcceleratorName": "H100", "GpuCount": 1, "EffectiveCapacity": 8192, "VLLMParams": dict(dp), "LearnedFrom": "deployment"})
    a.store.update("test-ns", "test-model", "variant-a", {"AcceleratorName": "H100", "GpuCount": 1, "TotalKvCapacityTokens": 50000, "EffectiveCapacity": 40000,
                                                          "VLLMParams": dict(dp), "LearnedFrom": "live"})
    r = a.analyze(inp([rm("pod-1", "variant-a", "H100", 10.0, 5000, 50000, 0, 500, 100)], [st("variant-a", 1), st("variant-b", 0)]))
    assert r["VariantCapacities"][1]["VariantName"] == "variant-b" and r["VariantCapacities"][1]["PerReplicaCapacity"] == 40000.0
    a = A()
    a.store.update("test-ns", "test-model", "variant-a", {"AcceleratorName": "A100", "GpuCount": 1, "EffectiveCapacity": 8192, "TotalKvCapacityTokens": 30000,
                                                          "VLLMParams": {"EffectiveMaxBatchedTokens": 8192, "MaxNumSeqs": 256, "NumGpuBlocksOverride": 1875, "BlockSize": 16},
                                                          "LearnedFrom": "deployment"})
    r = a.analyze(inp([rm("pod-1", "variant-x", "L40S", 5.0, 5000, 16000, 0, 500, 100)], [st("variant-x", 1), st("variant-a", 0)]))
    assert r["VariantCapacities"][1]["VariantName"] == "variant-a" and r["VariantCapacities"][1]["PerReplicaCapacity"] == 24000.0
    a = A()
    a.store.update("test-ns", "test-model", "variant-a", {"AcceleratorName": "H100", "GpuCount": 1, "EffectiveCapacity": 8192, "LearnedFrom": "deployment",
                                                          "VLLMParams": {"EffectiveMaxBatchedTokens": 8192, "MaxNumSeqs": 256}})
    r = a.analyze(inp([], [st("variant-a", 0)]))
    assert len(r["VariantCapacities"]) == 1 and r["VariantCapacities"][0]["PerReplicaCapacity"] == 8192.0
    # scaling signals (:480-536)
    r = A().analyze(inp([rm("pod-1", "variant-a", "H100", 10.0, 11000, 16000, 3, 100, 50)], one))
    assert r["RequiredCapacity"] > 0 and r["TotalDemand"] == 11000.0 + 3 * 100
    r = A().analyze(inp([rm("pod-1", "variant-a", "H100", 10.0, 1000, 16000, 0, 100, 50), rm("pod-2", "variant-a", "H100", 10.0, 1000, 16000, 0, 100, 50)],
                        [st("variant-a", 2)]))
    assert r["SpareCapacity"] > 0 and r["RequiredCapacity"] == 0.0
    r = A().analyze(inp([rm("pod-1", "variant-a", "H100", 10.0, 10000, 16000, 0, 100, 50)], one))
    assert r["RequiredCapacity"] == 0.0 and r["SpareCapacity"] == 0.0 and r["Utilization"] == 10000 / 12800
    # scheduler queue demand (:538-644)
    base = [rm("pod-1", "variant-a", "H100", 10.0, 5000, 16000, 0, 100, 50)]
    assert A().analyze(inp(base, one))["TotalDemand"] == 5000.0
    assert A().analyze(inp(base, one, {"QueueSize": 10, "QueueBytes": 8000}))["TotalDemand"] > 5000.0
    assert A().analyze(inp([rm("pod-1", "variant-a", "H100", 10.0, 5000, 16000, 0, 100, 50, hit=0.5)], one,
                           {"QueueSize": 10, "QueueBytes": 4000}))["TotalDemand"] == 6000.0
    assert A().analyze(inp(base, one, {"QueueSize": 10, "QueueBytes": 20000}))["TotalDemand"] == 10500.0
    # median (:646-667) through the variant aggregation: capacities {1,3,5,7}*1000/0.8 ... use k2 observed values
    caps = [1000, 3000, 5000, 7000]
    ms = [rm(f"pod-{i}", "variant-a", "H100", 10.0, c, 16000, 6, 100, 50) for i, c in enumerate(caps)]
    assert A().analyze(inp(ms, [st("variant-a", 4)]))["VariantCapacities"][0]["PerReplicaCapacity"] == 4000.0
    ms = [rm(f"pod-{i}", "variant-a", "H100", 10.0, c, 16000, 6, 100, 50) for i, c in enumerate([5000, 1000, 3000])]
    assert A().analyze(inp(ms, [st("variant-a", 3)]))["VariantCapacities"][0]["PerReplicaCapacity"] == 3000.0


def _vc(name, cost, count, cap, acc=""):
    return {"VariantName": name, "AcceleratorName": acc, "Cost": cost, "ReplicaCount": count, "PerReplicaCapacity": cap}


def _req(result, states, model="model-1", ns="default"):
    return {"ModelID": model, "Namespace": ns, "Result": result, "VariantStates": [st(n, c, p) for n, c, p in states]}


def _optimizer_cases(pkg, eng):
    opt = pkg.pipeline.CostAwareOptimizer(eng)
    assert opt.name() == "cost-aware"
    dm = lambda ds: {d["VariantName"]: d for d in ds}
    # scale-up (cost_aware_optimizer_test.go:30-142)
    d = dm(opt.optimize([_req({"RequiredCapacity": 5000, "VariantCapacities": [_vc("cheap", 5.0, 2, 10000, "A100"), _vc("expensive", 15.0, 1, 20000, "H100")]},
                              [("cheap", 2, 0), ("expensive", 1, 0)])]))
    assert d["cheap"]["TargetReplicas"] == 3 and d["expensive"]["TargetReplicas"] == 1
    d = dm(opt.optimize([_req({"RequiredCapacity": 5000, "VariantCapacities": [_vc("cheap", 5.0, 2, 10000), _vc("mid", 10.0, 1, 15000)]},
                              [("cheap", 2, 1), ("mid", 1, 0)])]))
    assert d["cheap"]["TargetReplicas"] == 3 and d["mid"]["TargetReplicas"] == 1
    d = dm(opt.optimize([_req({"RequiredCapacity": 5000, "VariantCapacities": [_vc("zero-cap", 1.0, 0, 0), _vc("normal", 10.0, 1, 10000)]},
                              [("zero-cap", 0, 0), ("normal", 1, 0)])]))
    assert d["normal"]["TargetReplicas"] == 2 and d["zero-cap"]["TargetReplicas"] == 0
    d = dm(opt.optimize([_req({"RequiredCapacity": 25000, "VariantCapacities": [_vc("cheap", 5.0, 1, 10000), _vc("mid", 10.0, 1, 15000)]},
                              [("cheap", 1, 0), ("mid", 1, 0)])]))
    assert d["cheap"]["TargetReplicas"] == 4 and d["mid"]["TargetReplicas"] == 1
    # scale-down (:144-262)
    d = dm(opt.optimize([_req({"SpareCapacity": 15000, "VariantCapacities": [_vc("cheap", 5.0, 3, 10000), _vc("expensive", 15.0, 2, 20000)]},
                              [("cheap", 3, 0), ("expensive", 2, 0)])]))
    assert d["expensive"]["TargetReplicas"] == 2 and d["cheap"]["TargetReplicas"] == 2
    d = dm(opt.optimize([_req({"SpareCapacity": 30000, "VariantCapacities": [_vc("expensive", 15.0, 1, 20000), _vc("cheap", 5.0, 1, 10000)]},
                              [("expensive", 1, 0), ("cheap", 1, 0)])]))
    assert d["expensive"]["TargetReplicas"] == 0 and d["cheap"]["TargetReplicas"] == 1
    d = dm(opt.optimize([_req({"SpareCapacity": 15000, "VariantCapacities": [_vc("expensive", 15.0, 1, 20000), _vc("cheap", 5.0, 1, 10000)]},
                              [("expensive", 1, 0), ("cheap", 1, 0)])]))
    assert d["expensive"]["TargetReplicas"] == 1 and d["cheap"]["TargetReplicas"] == 0
    d = dm(opt.optimize([_req({"SpareCapacity": 50000, "VariantCapacities": [_vc("expensive", 15.0, 2, 20000), _vc("mid", 10.0, 2, 15000), _vc("cheap", 5.0, 2, 10000)]},
                              [("expensive", 2, 0), ("mid", 2, 0), ("cheap", 2, 0)])]))
    assert (d["expensive"]["TargetReplicas"], d["mid"]["TargetReplicas"], d["cheap"]["TargetReplicas"]) == (0, 2, 1)
    # steady state, nil result, multi-model, metadata (:264-376)
    ds = opt.optimize([_req({"RequiredCapacity": 0, "SpareCapacity": 0, "VariantCapacities": [_vc("v1", 5.0, 2, 10000)]}, [("v1", 2, 0)])])
    assert len(ds) == 1 and ds[0]["Action"] == "no-change" and ds[0]["TargetReplicas"] == 2 and ds[0]["Reason"] == "V2 steady state"
    assert opt.optimize([_req(None, [])]) == []
    d = dm(opt.optimize([_req({"RequiredCapacity": 5000, "VariantCapacities": [_vc("m1-v1", 5.0, 1, 10000)]}, [("m1-v1", 1, 0)]),
                         _req({"SpareCapacity": 10000, "VariantCapacities": [_vc("m2-v1", 10.0, 2, 10000)]}, [("m2-v1", 2, 0)], model="model-2")]))
    assert (d["m1-v1"]["Action"], d["m1-v1"]["TargetReplicas"]) == ("scale-up", 2)
    assert (d["m2-v1"]["Action"], d["m2-v1"]["TargetReplicas"]) == ("scale-down", 1)
    assert d["m1-v1"]["Reason"] == "V2 scale-up (optimizer: cost-aware, required: 5000)"
    ds = opt.optimize([_req({"RequiredCapacity": 5000, "VariantCapacities": [_vc("v1", 5.0, 1, 10000, "A100")]}, [("v1", 1, 0)], ns="ns-1")])
    assert (ds[0]["ModelID"], ds[0]["Namespace"], ds[0]["AcceleratorName"], ds[0]["Cost"]) == ("model-1", "ns-1", "A100", 5.0)


def _enforcer_cases(pkg, eng):
    va = lambda *xs: [{"VariantName": n, "Cost": c} for n, c in xs]
    E = lambda f: pkg.pipeline.Enforcer(eng, f)
    boom = lambda *a: (_ for _ in ()).throw(RuntimeError("prometheus unavailable"))
    # scale-to-zero enabled (enforcer_test.go:30-146)
    t, app = E(lambda *a: 0).enforce_policy("m", "ns", {"variant-a": 2, "variant-b": 1}, va(("variant-a", 1.0), ("variant-b", 2.0)), True)
    assert app and t == {"variant-a": 0, "variant-b": 0}
    t, app = E(lambda *a: 10).enforce_policy("m", "ns", {"variant-a": 2, "variant-b": 1}, va(("variant-a", 1.0), ("variant-b", 2.0)), True)
    assert not app and t == {"variant-a": 2, "variant-b": 1}
    t, app = E(boom).enforce_policy("m", "ns", {"variant-a": 2, "variant-b": 1}, va(("variant-a", 1.0), ("variant-b", 2.0)), True)
    assert not app and t == {"variant-a": 2, "variant-b": 1}
    # scale-to-zero disabled (:148-222)
    t, app = E(lambda *a: 0).enforce_policy("m", "ns", {"variant-a": 0, "variant-b": 0}, va(("variant-a", 2.0), ("variant-b", 1.0)), False)
    assert app and t == {"variant-a": 0, "variant-b": 1}
    t, app = E(lambda *a: 0).enforce_policy("m", "ns", {"variant-a": 2, "variant-b": 0}, va(("variant-a", 2.0), ("variant-b", 1.0)), False)
    assert not app and t == {"variant-a": 2, "variant-b": 0}
    # tie -> alphabetical (:260-296); missing cost -> DefaultVariantCost 10 (:298-333)
    t, app = E(lambda *a: 0).enforce_policy("m", "ns", {"variant-z": 0, "variant-a": 0}, va(("variant-z", 1.0), ("variant-a", 1.0)), False)
    assert app and t == {"variant-a": 1, "variant-z": 0}
    t, app = E(lambda *a: 0).enforce_policy("m", "ns", {"variant-a": 0, "variant-missing": 0}, va(("variant-a", 100.0)), False)
    assert app and t == {"variant-a": 0, "variant-missing": 1}


# ---- CPU: oracle + host logic -------------------------------------------------------------------------------------------------------
def test_estimate_capacity_from_params(pkg, oracle):
    """analyzer_test.go:444-478 — the oracle and the caller-side chain of the mirror agree with the reference values."""
    f = pkg.pipeline.estimate_capacity_from_params
    for B, S, i, o, want in ((4096, 256, 500, 100, 140800), (8192, 64, 100, 200, 12800), (8192, 256, 500, 0, 0)):
        assert oracle.estimate_capacity_from_params(B, S, i, o) == want
        assert f({"EffectiveMaxBatchedTokens": B, "MaxNumSeqs": S}, i, o) == want
    assert f(None, 500, 100) == 0


def test_v2_analyzer_reference_cases_oracle(pkg, oracle):
    _analyzer_cases(pkg, OracleEngine(oracle))


def test_cost_aware_optimizer_reference_cases_oracle(pkg, oracle):
    _optimizer_cases(pkg, OracleEngine(oracle))


def test_enforcer_reference_cases_oracle(pkg, oracle):
    _enforcer_cases(pkg, OracleEngine(oracle))


# ---- seeded random batches ----------------------------------------------------------------------------------------------------------------
def random_v2_batch(M, seed, max_variants=6, max_replicas=9):
    g = np.random.default_rng(seed)
    nv = g.integers(0, max_variants + 1, M)
    mvo = np.concatenate([[0], np.cumsum(nv)]).astype(np.int32)
    V = int(mvo[-1])
    nr = g.integers(0, max_replicas + 1, V)
    vro = np.concatenate([[0], np.cumsum(nr)]).astype(np.int32)
    P = int(vro[-1])
    cap = g.choice([0, 8000, 16000, 32000, 64000], P, p=[0.1, 0.2, 0.3, 0.2, 0.2]).astype(np.int64)
    used = (cap * g.uniform(0, 1.1, P)).astype(np.int64)
    order = np.arange(P, dtype=np.int32)
    for m in range(M):                                   # shuffle every model's slice order
        a, b = vro[mvo[m]], vro[mvo[m + 1]]
        order[a:b] = g.permutation(np.arange(a, b))
    k2 = np.where(g.random(P) < 0.5, -1, (cap * g.uniform(0.2, 1.2, P)).astype(np.int64))
    return dict(n_models=M, n_variants=V, n_replicas=P, model_variant_off=mvo, variant_replica_off=vro,
                rep_total_kv_tokens=cap, rep_tokens_in_use=used, rep_queue_length=g.integers(0, 12, P).astype(np.int64),
                rep_avg_input_tokens=np.where(g.random(P) < 0.2, 0.0, g.uniform(10, 2000, P)),
                rep_avg_output_tokens=np.where(g.random(P) < 0.2, 0.0, g.uniform(10, 800, P)),
                rep_prefix_hit_rate=g.uniform(0, 0.9, P), rep_k2=k2, rep_slice_order=order,
                var_current=g.integers(0, 10, V).astype(np.int32), var_pending=g.integers(0, 3, V).astype(np.int32),
                var_fallback_capacity=np.where(g.random(V) < 0.5, 0.0, g.uniform(1000, 50000, V).round()),
                cfg_kv_threshold=g.choice([0.8, 0.85, 0.9], M), cfg_scale_up_threshold=g.choice([0.0, 0.85, 0.9], M),
                cfg_scale_down_boundary=g.choice([0.0, 0.6, 0.7], M),
                sched_queue_size=np.where(g.random(M) < 0.5, 0, g.integers(0, 50, M)).astype(np.int64),
                sched_queue_bytes=np.where(g.random(M) < 0.5, 0, g.integers(0, 200000, M)).astype(np.int64))


def random_optimizer_batch(M, seed, max_variants=40):
    g = np.random.default_rng(seed)
    nv = g.integers(0, max_variants + 1, M)
    mvo = np.concatenate([[0], np.cumsum(nv)]).astype(np.int32)
    V = int(mvo[-1])
    cap = np.where(g.random(V) < 0.15, 0.0, g.choice([4000.0, 8000.0, 12800.0, 20000.0], V))
    mode = g.integers(0, 3, M)
    return dict(model_variant_off=mvo, mod_required_capacity=np.where(mode == 0, g.uniform(1, 90000, M).round(), 0.0),
                mod_spare_capacity=np.where(mode == 1, g.uniform(1, 90000, M).round(), 0.0),
                mod_has_result=(g.random(M) < 0.9).astype(np.uint8), var_current=g.integers(0, 6, V).astype(np.int32),
                var_cost=g.choice([1.0, 2.5, 5.0, 5.0, 10.0, 15.0], V), var_per_replica_capacity=cap)


def random_enforcer_batch(M, seed, max_variants=40):
    g = np.random.default_rng(seed)
    nv = g.integers(0, max_variants + 1, M)
    mvo = np.concatenate([[0], np.cumsum(nv)]).astype(np.int32)
    V = int(mvo[-1])
    tgt = np.where(g.random(V) < 0.1, -1, g.integers(0, 4, V)).astype(np.int32)
    zero = g.random(M) < 0.5
    for m in np.where(zero)[0]:
        a, b = mvo[m], mvo[m + 1]
        tgt[a:b] = np.where(tgt[a:b] >= 0, 0, -1)
    return dict(model_variant_off=mvo, mod_scale_to_zero_enabled=(g.random(M) < 0.5).astype(np.uint8),
                mod_request_count=np.where(g.random(M) < 0.5, 0.0, g.uniform(0, 100, M)),
                mod_request_error=(g.random(M) < 0.1).astype(np.uint8), var_cost=g.choice([1.0, 2.0, 2.0, 5.0, 10.0, 20.0], V),
                var_has_cost=(g.random(V) < 0.85).astype(np.uint8), var_target=tgt)


def test_random_batches_oracle_invariants(oracle):
    """size-independent properties on the oracle itself: an optimizer never moves a model both ways, an enforced model
    without scale-to-zero keeps a replica, analyzer signals are non-negative and exclusive of nothing they should not be."""
    d = random_optimizer_batch(300, 11)
    t = oracle.cost_aware_optimize(d)
    mvo = d["model_variant_off"]
    for m in range(300):
        a, b = mvo[m], mvo[m + 1]
        if not d["mod_has_result"][m]:
            assert (t[a:b] == -1).all()
            continue
        delta = t[a:b] - d["var_current"][a:b]
        assert not ((delta > 0).any() and (delta < 0).any())
        if d["mod_required_capacity"][m] > 0:
            assert (delta >= 0).all()
        elif d["mod_spare_capacity"][m] > 0:
            assert (delta <= 0).all() and (t[a:b] >= 0).all()
    e = random_enforcer_batch(300, 12)
    t2, app = oracle.enforce(e)
    for m in range(300):
        a, b = e["model_variant_off"][m], e["model_variant_off"][m + 1]
        present = e["var_target"][a:b] >= 0
        if not e["mod_scale_to_zero_enabled"][m] and present.any():
            assert t2[a:b][present].sum() >= 1
    o = oracle.saturation_v2(random_v2_batch(200, 13))
    assert (o["mod_required_capacity"] >= 0).all() and (o["mod_spare_capacity"] >= 0).all()
    assert (o["rep_effective"] <= np.maximum(o["rep_k1"], 0)).all()


# ---- GPU -----------------------------------------------------------------------------------------------------------------------------------
def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint64) if a.dtype == np.float64 else a


@pytest.mark.gpu
@pytest.mark.parametrize("M,seed", [(1, 1), (50, 2), (3000, 3), (20000, 4)])
def test_saturation_v2_matches_oracle(engine, oracle, M, seed):
    d = random_v2_batch(M, seed)
    g, o = engine.saturation_v2(d), oracle.saturation_v2(d)
    for k in o:
        assert np.array_equal(_bits(g[k]), _bits(o[k])), k
    d2 = dict(d); d2["rep_slice_order"] = None; d2["sched_queue_size"] = None; d2["sched_queue_bytes"] = None
    g, o = engine.saturation_v2(d2), oracle.saturation_v2(d2)
    for k in o:
        assert np.array_equal(_bits(g[k]), _bits(o[k])), k


@pytest.mark.gpu
@pytest.mark.parametrize("max_variants,max_replicas,seed", [(40, 16, 11), (6, 120, 12), (70, 30, 13)])
def test_saturation_v2_staged_and_unstaged_models(engine, oracle, max_variants, max_replicas, seed):
    """Models of up to 256 replicas are staged through shared memory, larger ones read lane-per-variant from global
    memory (pipeline_v2_kernel.cuh V2_STAGE); both sides of the limit, variants above 8 replicas (the general median)
    and models wider than a warp in one batch."""
    d = random_v2_batch(300, seed, max_variants=max_variants, max_replicas=max_replicas)
    sizes = np.diff(d["variant_replica_off"][d["model_variant_off"]])
    assert (sizes <= 256).any() and (sizes > 256).any()
    g, o = engine.saturation_v2(d), oracle.saturation_v2(d)
    for k in o:
        assert np.array_equal(_bits(g[k]), _bits(o[k])), k


@pytest.mark.gpu
@pytest.mark.parametrize("M,seed", [(1, 5), (400, 6), (30000, 7)])
def test_cost_aware_and_enforcer_match_oracle(engine, oracle, M, seed):
    d = random_optimizer_batch(M, seed)
    assert np.array_equal(engine.cost_aware_optimize(d), oracle.cost_aware_optimize(d))
    e = random_enforcer_batch(M, seed + 100)
    gt, ga = engine.enforce(e)
    ot, oa = oracle.enforce(e)
    assert np.array_equal(gt, ot) and np.array_equal(ga, oa)


@pytest.mark.gpu
def test_v2_reference_cases_device(pkg, engine):
    _analyzer_cases(pkg, engine)
    _optimizer_cases(pkg, engine)
    _enforcer_cases(pkg, engine)


@pytest.mark.gpu
def test_v2_pipeline_end_to_end_device(pkg, engine, oracle):
    """analyzer -> optimizer -> enforcer -> limiter on one model, device stages chained through the mirrors, and the
    same chain on the oracle stand-in."""
    def run(eng):
        an = pkg.pipeline.SaturationAnalyzerV2(eng)
        res = an.analyze(inp([rm("p1", "a-cheap", "A100", 5.0, 12000, 16000, 2, 100, 50), rm("p2", "b-exp", "H100", 15.0, 30000, 32000, 1, 100, 50)],
                             [st("a-cheap", 1), st("b-exp", 1)]))
        ds = pkg.pipeline.CostAwareOptimizer(eng).optimize([{"ModelID": "test-model", "Namespace": "test-ns", "Result": res,
                                                             "VariantStates": [st("a-cheap", 1), st("b-exp", 1)]}])
        t, _ = pkg.pipeline.Enforcer(eng, lambda *a: 5).enforce_policy("test-model", "test-ns", {d["VariantName"]: d["TargetReplicas"] for d in ds},
                                                                       [{"VariantName": "a-cheap", "Cost": 5.0}], False)
        return res, ds, t
    r1, d1, t1 = run(engine)
    r2, d2, t2 = run(OracleEngine(oracle))
    assert r1 == r2 and d1 == d2 and t1 == t2
    assert r1["RequiredCapacity"] > 0 and t1["a-cheap"] >= 2


@pytest.mark.gpu
@pytest.mark.parametrize("M,seed", [(1, 21), (700, 22), (15000, 23)])
def test_fused_pipeline_equals_the_three_stages(engine, oracle, M, seed):
    """wva_pipeline_v2 (one upload, three chained launches, one download) == the composition of the three stage calls,
    on the device and on the oracle; variant names are NOT in index order (random name ranks)."""
    d = random_v2_batch(M, seed)
    g = np.random.default_rng(seed + 1000)
    V = int(d["n_variants"])
    cost = g.choice([1.0, 2.0, 2.0, 5.0, 10.0], V)
    rank = np.zeros(V, np.int32)
    mvo = d["model_variant_off"]
    for m in range(M):
        rank[mvo[m]:mvo[m + 1]] = g.permutation(mvo[m + 1] - mvo[m])
    s2z = (g.random(M) < 0.4).astype(np.uint8)
    cnt = np.where(g.random(M) < 0.5, 0.0, 3.0)
    err = (g.random(M) < 0.1).astype(np.uint8)
    go, gt, ga = engine.pipeline_v2(d, cost, s2z, cnt, err, rank)
    oo, ot, oa = oracle.pipeline_v2(d, cost, s2z, cnt, err, rank)
    assert np.array_equal(gt, ot) and np.array_equal(ga, oa)
    for k in oo:
        assert np.array_equal(_bits(go[k]), _bits(oo[k])), k
    # and against the separate device calls (name order == index order when no ranks are given)
    out = engine.saturation_v2(d)
    t = engine.cost_aware_optimize(dict(model_variant_off=mvo, mod_required_capacity=out["mod_required_capacity"],
                                        mod_spare_capacity=out["mod_spare_capacity"], var_current=d["var_current"], var_cost=cost,
                                        var_per_replica_capacity=out["var_per_replica_capacity"]))
    t2, a2 = engine.enforce(dict(model_variant_off=mvo, mod_scale_to_zero_enabled=s2z, mod_request_count=cnt, mod_request_error=err,
                                 var_cost=cost, var_has_cost=None, var_target=t))
    _, ft, fa = engine.pipeline_v2(d, cost, s2z, cnt, err, None)
    assert np.array_equal(ft, t2) and np.array_equal(fa, a2)


@pytest.mark.gpu
def test_v2_entry_points_reject_malformed_offsets(pkg, engine):
    d = random_v2_batch(20, 31)
    bad = dict(d); bad["variant_replica_off"] = d["variant_replica_off"].copy(); bad["variant_replica_off"][3] = 10**6
    with pytest.raises(pkg.WvaError):
        engine.saturation_v2(bad)
    bad = dict(d); bad["rep_slice_order"] = d["rep_slice_order"].copy()
    if bad["rep_slice_order"].size:
        bad["rep_slice_order"][0] = int(d["n_replicas"]) - 1 if d["variant_replica_off"][d["model_variant_off"][1]] < d["n_replicas"] - 1 else 0
        bad["rep_slice_order"][0] = int(d["n_replicas"]) + 5
        with pytest.raises(pkg.WvaError):
            engine.saturation_v2(bad)
    o = random_optimizer_batch(10, 32)
    o["model_variant_off"] = o["model_variant_off"].copy(); o["model_variant_off"][-1] += 1
    with pytest.raises(pkg.WvaError):
        engine.cost_aware_optimize(o)


def test_v2_analyzer_batch_equals_single_calls(pkg, oracle):
    """Every model of a cycle in one batched call == one Analyze per model (the k2 history and the store are keyed by
    model, so models do not interact); a second cycle sees the history of the first."""
    eng = OracleEngine(oracle)
    g = np.random.default_rng(17)

    def model(i, sat_queue):
        ms, sts = [], []
        for v in range(int(g.integers(1, 4))):
            n = int(g.integers(0, 4))
            for k in range(n):
                cap = int(g.choice([16000, 32000]))
                ms.append(rm(f"m{i}-v{v}-p{k}", f"v{v}", g.choice(["A100", "H100"]), float(g.choice([5.0, 10.0])), int(cap * g.uniform(0.1, 0.9)), cap,
                             int(g.integers(5, 9)) if sat_queue else int(g.integers(0, 4)), float(g.uniform(50, 900)), float(g.choice([40.0, 300.0, 700.0]))))
            sts.append(st(f"v{v}", n + int(g.integers(0, 2)), int(g.integers(0, 2))))
        d = inp(ms, sts, {"QueueSize": int(g.integers(0, 9)), "QueueBytes": int(g.integers(0, 9000))} if i % 2 else None)
        d["ModelID"] = f"model-{i}"
        return d

    cycle1 = [model(i, True) for i in range(12)]
    cycle2 = [dict(m, ReplicaMetrics=[dict(r, QueueLength=1) for r in m["ReplicaMetrics"]]) for m in cycle1]   # queues drained
    batch = pkg.pipeline.SaturationAnalyzerV2(eng)
    singles = pkg.pipeline.SaturationAnalyzerV2(eng)
    for cyc in (cycle1, cycle2):
        rb = batch.analyze_batch(cyc)
        rs = [singles.analyze(m) for m in cyc]
        assert rb == rs
    assert batch.history == singles.history and batch.history       # cycle 1 observed k2 values, cycle 2 used them
    assert batch.store.records == singles.store.records
