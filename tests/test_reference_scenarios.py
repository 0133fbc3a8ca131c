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
