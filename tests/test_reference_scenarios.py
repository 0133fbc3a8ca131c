"""The systems the reference's own tests build (tests/ref_scenarios.py), through the oracle on the CPU and through
`Manager` -> C-ABI on the GPU, asserting what the reference's tests assert and oracle == device bit for bit.

Reference tests mirrored: pkg/core/system_test.go (SetFromSpec :42, Calculate :1186, AllocateByType :1284,
GenerateSolution :1392); pkg/solver/greedy_test.go (:240-965); pkg/solver/solver_test.go SolveUnlimited (:280-425).
"""
import numpy as np
import pytest

from tests import ref_scenarios as rs

F32 = ("cost", "value", "itl", "ttft", "rho", "max_arrv_rate")
INTS = ("state", "acc", "num_replicas", "batch_size")


def _oracle_optimize(pkg, oracle, spec):
    d, idx = pkg.manager.flatten_spec(spec)
    cand = oracle.calculate(d)
    sol = oracle.solve(d, cand)
    return d, idx, cand, sol


# ---- flatten: System.SetFromSpec ----------------------------------------------------------------------------------------
def test_set_from_spec_counts_and_fields(pkg):
    """system_test.go:130-214: one accelerator / model / server / class / capacity entry, fields carried over."""
    d, idx = pkg.manager.flatten_spec(rs.single_a100(unlimited=False))
    assert (d["n_acc"], d["n_models"], d["n_servers"], d["n_types"]) == (1, 1, 1, 1)
    assert idx.acc == ["A100"] and idx.types == ["GPU_A100"] and idx.models == ["test-model"]
    assert d["acc_cost"][0] == np.float32(1.0) and d["acc_multiplicity"][0] == 1 and d["type_count"][0] == 4
    assert d["perf_max_batch"][0, 0] == 16 and d["perf_at_tokens"][0, 0] == 100 and d["perf_acc_count"][0, 0] == 1
    assert (d["perf_alpha"][0, 0], d["perf_beta"][0, 0], d["perf_gamma"][0, 0]) == (np.float32(10), np.float32(2), np.float32(0.1))
    assert d["srv_priority"][0] == 1 and d["srv_target_present"][0] == 1
    assert (d["srv_slo_itl"][0], d["srv_slo_ttft"][0], d["srv_slo_tps"][0]) == (100, 1000, 50)
    assert d["srv_arrival"][0] == 30 and d["srv_in_tokens"][0] == 100 and d["srv_out_tokens"][0] == 200
    assert d["unlimited"] is False and d["saturation_policy"] == "None" and d["delayed_best_effort"] is False


def test_flatten_defaults_and_unknowns(pkg):
    """server.go:38-41,92-97 and serviceclass.go:28-31: class "" -> "Free", unknown class -> lowest priority and no
    target, out-of-range priority -> default, unknown model -> -1, unknown current accelerator -> sentinel, a type
    named only by an accelerator has capacity 0, a repeated name replaces the earlier entry."""
    spec = rs.greedy_base()
    spec["serviceClassData"]["serviceClasses"].append({"name": "odd", "priority": 1000, "modelTargets": []})
    spec["serverData"]["servers"] += [
        rs._server("s-noclass", "llama-7b", "", 5, 10, 10),
        rs._server("s-badclass", "llama-7b", "nope", 5, 10, 10),
        rs._server("s-odd", "llama-7b", "odd", 5, 10, 10),
        rs._server("s-badmodel", "nope", "high-priority", 5, 10, 10, cur_acc="B300"),
        rs._server("server1", "llama-13b", "high-priority", 7, 10, 10, cur_acc="H100", cur_rep=3, cur_cost=6.0),
    ]
    spec["acceleratorData"]["accelerators"].append(rs._acc("L40", "GPU_L40", 0.5))
    d, idx = pkg.manager.flatten_spec(spec)
    i = idx.server_of
    assert idx.types == ["GPU_A100", "GPU_H100", "GPU_L40"] and d["type_count"].tolist() == [4, 2, 0]
    assert d["srv_priority"][i["s-noclass"]] == 100 and d["srv_target_present"][i["s-noclass"]] == 0
    assert d["srv_priority"][i["s-badclass"]] == 100 and d["srv_target_present"][i["s-badclass"]] == 0
    assert d["srv_priority"][i["s-odd"]] == 100
    assert d["srv_model"][i["s-badmodel"]] == -1 and d["srv_cur_acc"][i["s-badmodel"]] == pkg.manager.CUR_ACC_UNKNOWN
    assert d["n_servers"] == 3 + 4                                   # server1 replaced, not duplicated
    assert d["srv_model"][i["server1"]] == idx.model_of["llama-13b"] and d["srv_arrival"][i["server1"]] == 7
    assert d["srv_cur_acc"][i["server1"]] == idx.acc_of["H100"] and d["srv_cur_replicas"][i["server1"]] == 3
    assert d["srv_cur_acc"][i["server2"]] == pkg.manager.CUR_ACC_EMPTY
    assert d["perf_present"][idx.model_of["llama-7b"], idx.acc_of["L40"]] == 0


# ---- oracle on the reference's systems (CPU) ------------------------------------------------------------------------------
def test_system_calculate_oracle(pkg, oracle):
    """system_test.go:1248-1282: after Calculate the server has an A100 candidate with positive replicas and batch,
    non-negative cost."""
    d, idx, cand, _ = _oracle_optimize(pkg, oracle, rs.single_a100())
    assert cand["state"][0, 0] == 1 and cand["num_replicas"][0, 0] > 0 and cand["batch_size"][0, 0] > 0
    assert cand["cost"][0, 0] >= 0


def test_allocate_by_type_and_solution_oracle(pkg, oracle):
    """system_test.go:1352-1390 / :1460-1495: GPU_A100 entry with limit 4 and count, cost >= 0; the solution names
    the server with an accelerator and positive replicas."""
    spec = rs.single_a100(cur_acc="A100", cur_rep=2)
    d, idx, cand, sol = _oracle_optimize(pkg, oracle, spec)
    assert sol["state"][0] == 1 and sol["type_count"][0] == sol["num_replicas"][0] * 1 * 1 and sol["type_cost"][0] >= 0
    out = pkg.manager.solution_to_spec(sol, idx, {s["name"]: s for s in spec["serverData"]["servers"]})
    a = out["allocations"]["test-server"]
    assert a["accelerator"] == "A100" and a["numReplicas"] > 0 and a["load"]["arrivalRate"] == 30


@pytest.mark.parametrize("name", sorted(rs.greedy_scenarios()))
def test_greedy_scenarios_oracle(pkg, oracle, name):
    """greedy_test.go:240-965: the allocated-count window each test asserts; capacity is never exceeded."""
    spec, watch, lo, hi = rs.greedy_scenarios()[name]
    d, idx, cand, sol = _oracle_optimize(pkg, oracle, spec)
    n_alloc = sum(int(sol["state"][idx.server_of[s]] != 0) for s in watch)
    assert lo <= n_alloc <= hi, (name, n_alloc)
    assert (sol["type_count"] <= d["type_count"]).all()
    if name == "basic":
        assert (cand["state"][idx.server_of["server1"]] != 0).any()          # "should have candidate allocations"
    if name == "high_load":                                                  # llama-13b has no low-priority target
        assert (cand["state"][idx.server_of["server3"]] == 0).all()
    if name == "edge_cases":                                                 # zero load: the empty allocation
        z = idx.server_of["zero-load-server"]
        assert sol["state"][z] != 0 and sol["cost"][z] >= 0


def test_solve_unlimited_on_greedy_base_oracle(pkg, oracle):
    """solver_test.go:280-425: unlimited mode gives every server with candidates its minimum-value candidate."""
    spec = rs.greedy_base()
    spec["optimizerData"]["optimizer"]["unlimited"] = True
    d, idx, cand, sol = _oracle_optimize(pkg, oracle, spec)
    for i in range(d["n_servers"]):
        feas = cand["state"][i] != 0
        assert feas.any() and sol["state"][i] != 0
        assert sol["value"][i] == np.where(feas, cand["value"][i], np.inf).min()


class _OracleSizer:
    """engine stand-in for the CPU test of Manager.analyze_model: load/calculate/candidates answered by the oracle"""

    def __init__(self, oracle):
        self.o = oracle

    def load_system(self, d):
        self.d = d

    def calculate(self):
        self.c = self.o.calculate(self.d)

    def candidates(self):
        return self.c


def test_model_analyzer_adapter(pkg, oracle):
    """internal/modelanalyzer: one entry per accelerator with a non-nil allocation, QPS = float32(maxArrv * 1000),
    unknown server -> empty response (analyzer.go:24-33)."""
    m = pkg.manager.Manager(_OracleSizer(oracle), rs.greedy_base())
    r = m.analyze_model("server1")["Allocations"]
    assert set(r) == {"A100", "H100"}
    for acc, e in r.items():
        assert e["Reason"] == "markovian analysis" and e["RequiredPrefillQPS"] == e["RequiredDecodeQPS"]
        assert e["RequiredPrefillQPS"] == float(np.float32(e["Allocation"]["maxArrvRatePerReplica"]) * np.float32(1000))
        assert e["Allocation"]["accelerator"] == acc and e["Allocation"]["numReplicas"] >= 1
    assert m.analyze_model("nope") == {"Allocations": {}}
    s = rs.greedy_base()
    s["serverData"]["servers"].append(rs._server("lonely", "llama-13b", "low-priority", 5, 10, 10))   # no target for the class
    assert pkg.manager.Manager(_OracleSizer(oracle), s).analyze_model("lonely") == {"Allocations": {}}


# ---- the same systems through Manager -> C-ABI (GPU) ---------------------------------------------------------------------------
def _device_vs_oracle(pkg, engine, oracle, spec):
    d, idx, cand, want = _oracle_optimize(pkg, oracle, spec)
    m = pkg.manager.Manager(engine, spec)
    got_spec = m.optimize()
    got = m.last_solution
    gc = engine.candidates()
    for k in ("state", "num_replicas", "batch_size"):
        assert np.array_equal(gc[k], cand[k]), k
    for k in F32:
        assert np.array_equal(gc[k].view(np.uint32), cand[k].view(np.uint32)), k
    for k in INTS:
        assert np.array_equal(got[k], want[k]), k
    for k in F32:
        assert np.array_equal(got[k].view(np.uint32), want[k].view(np.uint32)), k
    assert np.array_equal(got["type_count"], want["type_count"])
    servers = {s["name"]: s for s in m.spec["serverData"]["servers"]}
    assert got_spec == pkg.manager.solution_to_spec(want, idx, servers)
    return m, got_spec, idx


@pytest.mark.gpu
def test_model_analyzer_adapter_device(pkg, engine, oracle):
    a = pkg.manager.Manager(engine, rs.greedy_base()).analyze_model("server2")
    b = pkg.manager.Manager(_OracleSizer(oracle), rs.greedy_base()).analyze_model("server2")
    assert a == b and set(a["Allocations"]) == {"A100", "H100"}


@pytest.mark.gpu
def test_single_a100_device(pkg, engine, oracle):
    m, out, idx = _device_vs_oracle(pkg, engine, oracle, rs.single_a100(cur_acc="A100", cur_rep=2))
    a = out["allocations"]["test-server"]
    assert a["accelerator"] == "A100" and a["numReplicas"] > 0 and a["maxBatch"] > 0
    by_type = m.allocation_by_type()
    assert by_type["GPU_A100"]["limit"] == 4 and by_type["GPU_A100"]["count"] == a["numReplicas"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(rs.greedy_scenarios()))
def test_greedy_scenarios_device(pkg, engine, oracle, name):
    spec, watch, lo, hi = rs.greedy_scenarios()[name]
    m, out, idx = _device_vs_oracle(pkg, engine, oracle, spec)
    assert lo <= sum(1 for s in watch if s in out["allocations"]) <= hi
    for t, v in m.allocation_by_type().items():
        assert v["count"] <= v["limit"]


@pytest.mark.gpu
@pytest.mark.parametrize("policy", ["None", "PriorityExhaustive", "PriorityRoundRobin", "RoundRobin"])
@pytest.mark.parametrize("delayed", [False, True])
def test_greedy_base_all_policies_device(pkg, engine, oracle, policy, delayed):
    """Every SaturationPolicy x DelayedBestEffort on the reference's base system and on its scarce variant."""
    for caps in (None, {"GPU_A100": 1, "GPU_H100": 1}, {"GPU_A100": 0, "GPU_H100": 3}):
        _device_vs_oracle(pkg, engine, oracle, rs._with(rs.greedy_base(), policy=policy, delayed=delayed, caps=caps))


@pytest.mark.gpu
def test_unlimited_base_device(pkg, engine, oracle):
    spec = rs.greedy_base()
    spec["optimizerData"]["optimizer"]["unlimited"] = True
    _, out, _ = _device_vs_oracle(pkg, engine, oracle, spec)
    assert set(out["allocations"]) == {"server1", "server2", "server3"}


# ---- the pipeline interfaces: SaturationAnalyzer and Limiter, records with the reference's field names ------------------------
CFG = {"KvCacheThreshold": 0.8, "QueueLengthThreshold": 5, "KvSpareTrigger": 0.1, "QueueSpareTrigger": 3}


def _rm(variant, kvs, queues, cost=10.0, acc="A100"):
    return [{"PodName": f"{variant}-pod-{i}", "VariantName": variant, "ModelID": "test-model", "Namespace": "test-ns",
             "AcceleratorName": acc, "Cost": cost, "KvCacheUsage": kv, "QueueLength": q}
            for i, (kv, q) in enumerate(zip(kvs, queues))]


def _states(rows):
    return [{"VariantName": v, "CurrentReplicas": c, "DesiredReplicas": d, "PendingReplicas": p} for v, c, d, p in rows]


@pytest.mark.gpu
def test_analyze_model_saturation_reference_cases(pkg, engine):
    """internal/saturation/analyzer_test.go:17-323, through the interface the engine loop calls."""
    an = pkg.pipeline.SaturationAnalyzer(engine)
    # :17-60 scale-up on KV (spare 0.05 / 0.04 < 0.1), :62-100 on queue (spare 2 < 3), :102-140 no trigger
    a = an.analyze_model_saturation("test-model", "test-ns", _rm("v1", [.75, .76], [2, 2]), CFG)
    assert a["ShouldScaleUp"] and a["ScaleUpReason"].startswith("KV spare Saturation low (0.045 < 0.100)")
    assert a["TotalReplicas"] == 2 and a["NonSaturatedCount"] == 2
    a = an.analyze_model_saturation("test-model", "test-ns", _rm("v1", [.5, .5], [3, 3]), CFG)
    assert a["ShouldScaleUp"] and a["ScaleUpReason"] == "queue spare Saturation low (2.0 < 3.0)"
    a = an.analyze_model_saturation("test-model", "test-ns", _rm("v1", [.5, .5], [1, 1]), CFG)
    assert not a["ShouldScaleUp"] and a["ScaleUpReason"] == ""
    # :142-192 two variants aggregate; :228-276 saturated replicas are named
    a = an.analyze_model_saturation("test-model", "test-ns", _rm("v1", [.70, .75], [2, 3]) + _rm("v2", [.60, .65], [1, 2], acc="H100"), CFG)
    assert a["TotalReplicas"] == 4 and a["NonSaturatedCount"] == 4 and len(a["VariantAnalyses"]) == 2
    assert [v["VariantName"] for v in a["VariantAnalyses"]] == ["v1", "v2"] and a["VariantAnalyses"][1]["AcceleratorName"] == "H100"
    a = an.analyze_model_saturation("test-model", "test-ns", _rm("v1", [.85, .50, .60], [2, 6, 2]), CFG)
    va = a["VariantAnalyses"][0]
    assert va["SaturatedReplicas"] == ["v1-pod-0", "v1-pod-1"] and va["NonSaturatedCount"] == 1
    assert va["MaxKvCacheUsage"] == .85 and va["MaxQueueLength"] == 6
    # :194-226 scale-down safety; :278-323 no metrics
    assert an.analyze_model_saturation("m", "n", _rm("v1", [.2, .3, .25], [1, 1, 1]), CFG)["ScaleDownSafe"]
    assert not an.analyze_model_saturation("m", "n", _rm("v1", [.7, .75], [2, 2]), CFG)["ScaleDownSafe"]
    assert not an.analyze_model_saturation("m", "n", _rm("v1", [.5], [2]), CFG)["ScaleDownSafe"]
    e = an.analyze_model_saturation("m", "n", [], CFG)
    assert e["TotalReplicas"] == 0 and not e["ShouldScaleUp"] and not e["ScaleDownSafe"] and e["VariantAnalyses"] == []
    assert an.calculate_saturation_targets(e, _states([("v1", 3, 0, 0)])) == {"v1": 3}


@pytest.mark.gpu
def test_calculate_saturation_targets_reference_cases(pkg, engine):
    """analyzer_test.go:367-509: cheapest +1, most expensive -1, model-level transition blocking, metrics mismatch."""
    an = pkg.pipeline.SaturationAnalyzer(engine)
    three = lambda kv, q: (_rm("v1-expensive", [kv] * 2, [q] * 2, 20) + _rm("v2-cheap", [kv] * 2, [q] * 2, 5)
                           + _rm("v3-medium", [kv] * 2, [q] * 2, 15))
    two = lambda kv, q: _rm("v1-expensive", [kv] * 2, [q] * 2, 20) + _rm("v2-cheap", [kv] * 2, [q] * 2, 5)
    up = an.analyze_model_saturation("test-model", "test-ns", three(.75, 2), CFG)
    assert up["ShouldScaleUp"]
    s3 = _states([("v1-expensive", 2, 0, 0), ("v2-cheap", 2, 0, 0), ("v3-medium", 2, 0, 0)])
    assert an.calculate_saturation_targets(up, s3) == {"v1-expensive": 2, "v2-cheap": 3, "v3-medium": 2}
    down = an.analyze_model_saturation("test-model", "test-ns", three(.2, 1), CFG)
    assert not down["ShouldScaleUp"] and down["ScaleDownSafe"]
    assert an.calculate_saturation_targets(down, s3) == {"v1-expensive": 1, "v2-cheap": 2, "v3-medium": 2}
    up2 = an.analyze_model_saturation("test-model", "test-ns", two(.75, 2), CFG)
    assert an.calculate_saturation_targets(up2, _states([("v1-expensive", 2, 4, 0), ("v2-cheap", 2, 0, 0)])) == \
        {"v1-expensive": 4, "v2-cheap": 2}
    assert an.calculate_saturation_targets(up2, _states([("v1-expensive", 3, 0, 0), ("v2-cheap", 2, 0, 0)])) == \
        {"v1-expensive": 3, "v2-cheap": 2}
    # pending replicas on the cheapest variant: the next cheapest takes the replica (analyzer.go:363-392)
    s3p = _states([("v1-expensive", 2, 0, 0), ("v2-cheap", 2, 0, 1), ("v3-medium", 2, 0, 0)])
    assert an.calculate_saturation_targets(up, s3p) == {"v1-expensive": 2, "v2-cheap": 2, "v3-medium": 3}


@pytest.mark.gpu
def test_analyze_batch_equals_single_calls(pkg, engine, oracle):
    """Every model of a cycle in one launch == one call per model; == the oracle on the packed batch."""
    an = pkg.pipeline.SaturationAnalyzer(engine)
    rng = np.random.default_rng(5)
    models = []
    for m in range(40):
        rm, st = [], []
        for v in range(int(rng.integers(1, 5))):
            n = int(rng.integers(1, 6))
            rm += _rm(f"m{m}-v{v}", rng.uniform(0, 1, n).round(3).tolist(), rng.integers(0, 8, n).tolist(),
                      cost=float(rng.choice([5, 10, 20])))
            st.append((f"m{m}-v{v}", n + int(rng.integers(0, 2)), int(rng.choice([0, 0, n])), int(rng.integers(0, 2))))
        models.append({"modelID": f"m{m}", "namespace": "ns", "replicaMetrics": rm, "config": CFG, "variantStates": _states(st)})
    batch = an.analyze_batch(models)
    for m, (a, t) in zip(models, batch):
        a1 = an.analyze_model_saturation(m["modelID"], "ns", m["replicaMetrics"], CFG)
        t1 = an.calculate_saturation_targets(a1, m["variantStates"])
        a1.pop("_src")
        assert a == a1 and t == t1
    d, _, _ = an._pack([(m["replicaMetrics"], m["config"], m["variantStates"]) for m in models])
    want = oracle.saturation_v1(d)
    flat = [t[v] for m, (a, t) in zip(models, batch) for v in sorted(t)]
    assert flat == [int(x) for x in want["var_target"] if x >= 0]


def _dec(name, acc, cur, tgt, gpr=2, spare=0.1, cost=10.0):
    return {"VariantName": name, "AcceleratorName": acc, "CurrentReplicas": cur, "TargetReplicas": tgt,
            "GPUsPerReplica": gpr, "SpareCapacity": spare, "Cost": cost, "Action": "scale-up"}


@pytest.mark.gpu
def test_limiter_reference_cases(pkg, engine):
    """default_limiter_test.go:152-335 and greedy_saturation_algorithm_test.go:54-270 through `Limiter.limit`."""
    lim = pkg.pipeline.Limiter(engine, "gpu-limiter", {"A100": 12})
    ds = [_dec("a", "A100", 1, 2, spare=.3), _dec("b", "A100", 1, 2, spare=.05), _dec("c", "A100", 1, 2, spare=.5)]
    lim.limit(ds)                                             # 6 GPUs in use, 6 free: everyone gets its replica
    assert [d["GPUsAllocated"] for d in ds] == [2, 2, 2] and not any(d["WasLimited"] for d in ds)
    assert ds[0]["DecisionSteps"][-1]["Reason"] == "allocated 2 GPUs for +1 replicas" and "LimitedBy" not in ds[0]
    lim = pkg.pipeline.Limiter(engine, "gpu-limiter", lambda: {"A100": 10})
    ds = [_dec("a", "A100", 1, 2, spare=.3), _dec("b", "A100", 1, 2, spare=.05), _dec("c", "A100", 1, 2, spare=.5)]
    lim.limit(ds)                                             # 4 free: most saturated first (.05, .3), .5 is limited
    assert [d["TargetReplicas"] for d in ds] == [2, 2, 1] and [d["WasLimited"] for d in ds] == [False, False, True]
    assert ds[2]["LimitedBy"] == "gpu-limiter" and ds[2]["DecisionSteps"][-1]["WasConstrained"]
    assert ds[2]["DecisionSteps"][-1]["Reason"] == "no scale-up (target=1, current=1)"
    lim = pkg.pipeline.Limiter(engine, "gpu-limiter", {"A100": 5})
    ds = [_dec("a", "A100", 1, 3, spare=.1, cost=5.0)]
    lim.limit(ds)                                             # 3 free, 2 per replica: one replica
    assert ds[0]["TargetReplicas"] == 2 and ds[0]["GPUsAllocated"] == 2 and ds[0]["WasLimited"]
    assert ds[0]["DecisionSteps"][-1]["Reason"] == "limited: allocated 2 GPUs for +1 replicas"
    lim = pkg.pipeline.Limiter(engine, "gpu-limiter", {"A100": 12, "H100": 0})   # 10 in use, 2 free
    ds = [_dec("exp", "A100", 1, 2, spare=.2, cost=20.0), _dec("cheap", "A100", 1, 2, spare=.2, cost=5.0),
          _dec("h", "H100", 0, 2), _dec("nopool", "L40", 1, 2), _dec("noacc", "", 1, 2), _dec("down", "A100", 3, 1)]
    lim.limit(ds)
    assert [d["TargetReplicas"] for d in ds[:2]] == [1, 2]    # equal spare: the cheaper variant first
    assert ds[2]["TargetReplicas"] == 0 and ds[3]["TargetReplicas"] == 1 and ds[4]["TargetReplicas"] == 1
    assert ds[5]["TargetReplicas"] == 1 and not ds[5]["WasLimited"]   # scale-down passes through
    pkg.pipeline.Limiter(engine, "gpu-limiter", {"A100": 1}).limit([])   # empty: no inventory refresh, no launch
