"""Pins the oracle (oracle/*.hpp) to every known-answer vector the reference's own tests hold for the
hot path.  Each test names the reference test it restates (paths under /root/reference).
The reference is Go and cannot run here; these vectors are what anchors the C++ restatement.
"""
import math

import numpy as np
import pytest


# ---- pkg/analyzer/queueanalyzer_test.go:197-247 TestPrefillParms_PrefillTime ---------------------------
@pytest.mark.parametrize("in_tok,batch,expected", [(0, 4.0, 0.0), (1000, 1.0, 32.0), (2000, 8.0, 208.0),
                                                   (500, 2.5, 29.25)])
def test_prefill_time(oracle, in_tok, batch, expected):
    assert abs(oracle.prefill_time(10.0, 0.01, 0.001, in_tok, 0, batch) - expected) <= 1e-6


# ---- queueanalyzer_test.go:249-295 TestDecodeParms_DecodeTime ---------------------------------------------
@pytest.mark.parametrize("batch,expected", [(1.0, 1.23), (4.0, 1.575), (8.0, 2.035), (2.5, 1.4025)])
def test_decode_time(oracle, batch, expected):
    assert abs(oracle.decode_time(1.0, 0.1, 0.01, 1, 1, batch) - expected) <= 1e-6


# ---- pkg/analyzer/utils_test.go:9-70 TestWithinTolerance ----------------------------------------------------
@pytest.mark.parametrize("x,v,tol,exp", [(1.0, 1.0, 0.01, True), (1.005, 1.0, 0.01, True), (1.02, 1.0, 0.01, False),
                                         (0.1, 0.0, 0.01, False), (1.0, 1.0, -0.01, True), (0.0, 0.0, 0.01, True)])
def test_within_tolerance(oracle, x, v, tol, exp):
    assert oracle.within_tolerance(x, v, tol) is exp


# ---- utils_test.go:72-223 TestBinarySearch --------------------------------------------------------------------
def test_binary_search_table(oracle):
    rc, x, ind = oracle.binary_search_poly(0.0, 10.0, 4.0, c2=1.0)            # find square root
    assert rc == 0 and ind == 0 and abs(x * x - 4.0) <= 0.1
    rc, x, ind = oracle.binary_search_poly(1.0, 5.0, 6.0, c1=2.0)             # linear, target in range
    assert rc == 0 and ind == 0 and abs(2 * x - 6.0) <= 0.1
    rc, x, ind = oracle.binary_search_poly(2.0, 5.0, 1.0, c1=2.0)             # below range
    assert rc == 0 and ind == -1 and x == 2.0
    rc, x, ind = oracle.binary_search_poly(1.0, 3.0, 10.0, c1=2.0)            # above range
    assert rc == 0 and ind == 1 and x == 3.0
    rc, x, ind = oracle.binary_search_poly(1.0, 5.0, -3.0, c1=-1.0)           # decreasing function
    assert rc == 0 and ind == 0 and abs(-x - -3.0) <= 0.1
    rc, _, _ = oracle.binary_search_poly(5.0, 1.0, 3.0, c1=2.0)               # invalid range -> error
    assert rc != 0
    rc, _, _ = oracle.binary_search_poly(4.0, 6.0, 5.0, c1=1.0, fail_at=5.0000001)  # eval error -> error
    assert rc != 0
    rc, x, ind = oracle.binary_search_poly(1.0, 5.0, 2.0, c1=2.0)             # target at boundary
    assert rc == 0 and ind == 0 and x == 1.0


# ---- utils_test.go:225-289 TestBinarySearch_EdgeCases ----------------------------------------------------------
def test_binary_search_edge_cases(oracle):
    assert oracle.binary_search_poly(1.0, 10.0, 5.0, c0=5.0)[0] == 0          # constant, matches
    assert oracle.binary_search_poly(1.0, 10.0, 3.0, c0=5.0)[0] == 0          # constant, no match: no error
    assert oracle.binary_search_poly(3.0, 3.0, 6.0, c1=2.0)[0] == 0           # zero range


# ---- pkg/analyzer/queuemodel_test.go:9-102 MM1K validity gate ---------------------------------------------------
@pytest.mark.parametrize("lam,mu,valid", [(1, 2, True), (0, 2, True), (-1, 2, False), (1, 0, False), (1, -1, False),
                                          (9.9, 1, True), (11, 1, False)])
def test_mm1k_validity(oracle, lam, mu, valid):
    st, p = oracle.mm1k_solve(10, lam, mu)
    assert bool(st["valid"]) is valid
    if valid and lam > 0:
        assert st["avgRespTime"] >= 0 and st["avgWaitTime"] >= 0 and st["throughput"] >= 0
        assert abs(p.sum() - 1.0) <= 1e-6                                    # queuemodel_test.go:152-222


# ---- queuemodel_test.go:325-400 state-dependent validity -------------------------------------------------------------
def test_statedep_validity(oracle):
    rc, rows, p = oracle.statedep_solve(5, [1.0, 2.0, 3.0], [0.5, 1.5, 2.8, 0.0])
    assert all(r["valid"] == 1.0 for r in rows)
    rc, rows, _ = oracle.statedep_solve(5, [1.0, 2.0, 3.0], [-1.0])
    assert rows[0]["valid"] == 0.0


# ---- queuemodel_test.go:461-533 sum p = 1, Little's law, MM1K == state-dependent with constant rate -----------------
def test_littles_law_and_equivalence(oracle):
    for lam in (0.5, 1.5, 2.5):
        rc, rows, p = oracle.statedep_solve(5, [3.0, 3.0, 3.0], [lam])
        r = rows[0]
        assert abs(p.sum() - 1.0) <= 1e-6
        assert abs(r["avgNumInSystem"] - r["throughput"] * r["avgRespTime"]) <= 1e-4
    st, _ = oracle.mm1k_solve(5, 1.5, 3.0)
    _, rows, _ = oracle.statedep_solve(5, [3.0], [1.5])
    assert abs(st["avgNumInSystem"] - rows[0]["avgNumInSystem"]) <= 1e-3
    assert abs(st["throughput"] - rows[0]["throughput"]) <= 1e-3


# ---- quirk Q1: the validity gate reads p[0] left by the previous solve ---------------------------------------------------
def test_stale_p0_gate(oracle):
    # K == 1: rho = 1 - float32(0) = 1 >= K -> invalid on the first (and every) call
    _, rows, _ = oracle.statedep_solve(1, [1.0], [0.5, 0.5])
    assert rows[0]["valid"] == 0.0 and rows[1]["valid"] == 0.0
    _, rows, _ = oracle.statedep_solve(2, [1.0], [0.5, 0.5])
    assert rows[0]["valid"] == 1.0 and rows[0]["rho"] == rows[1]["rho"]


# ---- queueanalyzer_test.go:337-534 Analyze / Size error conditions --------------------------------------------------------
CFG = dict(max_batch=8, max_queue=16, a=1.0, b=0.01, g=0.001, i=100, o=10)


def test_analyze_errors(oracle):
    rc, m, (rmin, rmax) = oracle.queue_analyze(rate=1.0, **CFG)
    assert rc == 0 and rmin > 0 and rmax > rmin
    for k in ("Throughput", "AvgRespTime", "AvgWaitTime", "AvgNumInServ", "AvgPrefillTime", "AvgTokenTime"):
        assert m[k] >= 0
    assert 0 <= m["Rho"] <= 1
    assert oracle.queue_analyze(rate=0.0, **CFG)[0] == 2
    assert oracle.queue_analyze(rate=-1.0, **CFG)[0] == 2
    assert oracle.queue_analyze(rate=rmax * 1.01, **CFG)[0] == 2
    assert oracle.queue_analyze(rate=rmax, **CFG)[0] == 0


def test_size_errors_and_ranges(oracle):
    assert oracle.queue_size(ttft=-1.0, itl=10.0, tps=0.0, **CFG)[0] == 2      # negative targets -> error
    assert oracle.queue_size(ttft=10.0, itl=-1.0, tps=0.0, **CFG)[0] == 2
    assert oracle.queue_size(ttft=10.0, itl=10.0, tps=-1.0, **CFG)[0] == 2
    rc, rates, m, ach, ns = oracle.queue_size(ttft=100.0, itl=5.0, tps=0.0, **CFG)
    assert rc == 0 and all(r > 0 for r in rates) and m["Throughput"] > 0
    rc, rates, m, ach, ns = oracle.queue_size(ttft=0.0, itl=0.0, tps=100.0, **CFG)
    assert rc == 0 and abs(rates[2] - 0.9 * rates[0]) <= 1e-3 * rates[0] and ns == 1
    bad = dict(CFG); bad["max_batch"] = 0
    assert oracle.queue_size(ttft=1.0, itl=1.0, tps=0.0, **bad)[0] == 1        # invalid configuration


# ---- pkg/core/allocation_test.go:235-284 TestAllocation_TransitionPenalty (exact compare) ----------------------------------
def test_transition_penalty(oracle):
    f32 = np.float32
    assert oracle.transition_penalty(0, 2, 100.0, 1, 0, 2, 100.0) == 0.0                    # same acc, same replicas
    assert oracle.transition_penalty(0, 2, 100.0, 1, 0, 3, 150.0) == 50.0                   # same acc, different replicas
    expected = float(f32(0.1) * (f32(100.0) + f32(120.0)) + (f32(120.0) - f32(100.0)))     # different accelerator
    assert oracle.transition_penalty(0, 2, 100.0, 1, 1, 2, 120.0) == expected


# ---- allocation_test.go:79-137,190-233,968-1126 zero-load allocation + Saturated ------------------------------------------------
def _one_server_system(pkg, arrival, ttft, itl, tps=0.0, min_rep=1, cost=100.0, alpha=5.0, beta=0.2, gamma=0.015,
                       max_batch=16, in_tok=100, out_tok=200):
    d = pkg.synth.queue_system(1, 1, max_batch, n_classes=1, stream=99)
    d["acc_cost"][:] = cost
    d["perf_alpha"][:] = alpha; d["perf_beta"][:] = beta; d["perf_gamma"][:] = gamma
    d["perf_acc_count"][:] = 1
    d["perf_at_tokens"][:] = out_tok
    d["srv_arrival"][:] = arrival
    d["srv_in_tokens"][:] = in_tok; d["srv_out_tokens"][:] = out_tok
    d["srv_slo_ttft"][:] = ttft; d["srv_slo_itl"][:] = itl; d["srv_slo_tps"][:] = tps
    d["srv_min_replicas"][:] = min_rep
    return d


def test_zero_load_allocation(pkg, oracle):
    c = oracle.calculate(_one_server_system(pkg, 0.0, 2000.0, 500.0))
    assert c["state"][0, 0] == 1 and c["num_replicas"][0, 0] == 1 and c["batch_size"][0, 0] == 16
    assert c["cost"][0, 0] == 100.0
    # Saturated(totalRate) = totalRate > replicas * maxArrvRatePerReplica*1000*60: false at 15000, true at 78132
    max_rpm = float(c["max_arrv_rate"][0, 0]) * 1000 * 60
    assert not (15000 > max_rpm) and (78132 > max_rpm)
    assert abs(max_rpm - 71642) < 1.0
    c = oracle.calculate(_one_server_system(pkg, 0.0, 2000.0, 500.0, min_rep=0))
    assert c["state"][0, 0] == 2 and c["num_replicas"][0, 0] == 0 and c["cost"][0, 0] == 0.0


# ---- allocation_test.go:576-773 CreateAllocation feasibility -------------------------------------------------------------------------
def test_create_allocation_feasibility(pkg, oracle):
    c = oracle.calculate(_one_server_system(pkg, 1200.0, 1.0, 0.1))          # unattainable targets -> nil
    assert c["state"][0, 0] == 0
    for arrival in (60.0, 120.0):
        c = oracle.calculate(_one_server_system(pkg, arrival, 2000.0, 500.0))
        assert c["state"][0, 0] == 1 and c["num_replicas"][0, 0] > 0
    c = oracle.calculate(_one_server_system(pkg, 60.0, 2000.0, 500.0, tps=2.0))
    assert c["state"][0, 0] == 1 and c["num_replicas"][0, 0] > 0


# ---- survey anchors: an independent numpy-float32 emulation of the same path (SURVEY.md §8c "Net") ------------------------------------
@pytest.mark.parametrize("a,b,g,N,I,O,ttft,itl,rpm,solves,rate_star,replicas", [
    (10, .2, .01, 16, 100, 200, 2000, 400, 1800, 24, 1.4361, 21),
    (8, .15, .008, 32, 100, 200, 2000, 400, 3000, 28, 2.1706, 24),
    (5, .2, .015, 16, 100, 200, 2000, 500, 120, 28, 1.1536, 2),
    (6.973, .027, .001, 64, 512, 128, 500, 24, 6000, 125, 7.1168, 15)])
def test_survey_anchors(oracle, a, b, g, N, I, O, ttft, itl, rpm, solves, rate_star, replicas):
    rc, rates, m, ach, ns = oracle.queue_size(N, 10 * N, a, b, g, I, O, ttft, itl, 0.0)
    assert rc == 0 and ns == solves
    assert abs(m["Throughput"] - rate_star) < 1e-4
    assert math.ceil(float(np.float32(rpm) / np.float32(60)) / m["Throughput"]) == replicas


# ---- pkg/solver: SolveUnlimited picks the minimum value; greedy structural properties ------------------------------------------------------
def test_solve_unlimited_min_value(pkg, oracle):
    d = pkg.synth.queue_system(30, 6, 16, stream=8)
    c = oracle.calculate(d)
    s = oracle.solve(d, c)
    for i in range(30):
        feas = c["state"][i] != 0
        if not feas.any():
            assert s["state"][i] == 0
            continue
        vals = np.where(feas, c["value"][i], np.inf)
        assert s["acc"][i] == int(np.argmin(vals)) or c["state"][i, int(np.argmin(vals))] == 2
        assert s["value"][i] == vals.min()


def test_greedy_structure(pkg, oracle):
    """greedy_test.go: resources are never over-committed; priority groups are served in order;
    when capacity is ample greedy == unlimited; exhausting every candidate leaves the server unallocated."""
    d = pkg.synth.queue_system(60, 6, 16, stream=9)
    c = oracle.calculate(d)
    un = oracle.solve(d, c)
    ample = pkg.synth.limit_capacity(d, un["type_count"] * 10, 1.0)
    g = oracle.solve(ample, c)
    assert np.array_equal(g["acc"], un["acc"]) and np.array_equal(g["num_replicas"], un["num_replicas"])
    for pol in ("None", "PriorityExhaustive", "PriorityRoundRobin", "RoundRobin"):
        for delayed in (False, True):
            lim = pkg.synth.limit_capacity(d, un["type_count"], 0.5)
            lim["saturation_policy"] = pol
            lim["delayed_best_effort"] = delayed
            g = oracle.solve(lim, c)
            assert (g["type_count"] <= lim["type_count"]).all(), (pol, delayed)
            assert (g["state"] == 0).sum() >= (un["state"] == 0).sum()
    zero = pkg.synth.limit_capacity(d, un["type_count"], 0.0)
    zero["type_count"][:] = 0
    g = oracle.solve(zero, c)
    assert ((g["state"] == 1) & (g["num_replicas"] > 0)).sum() == 0


# ---- internal/saturation/analyzer_test.go:17-140 scale-up / scale-down-safe flags ------------------------------------------------------------
def _sat_single(kvs, queues):
    n = len(kvs)
    return dict(n_models=1, n_variants=1, n_replicas=n, model_variant_off=[0, 1], variant_replica_off=[0, n],
                rep_kv=kvs, rep_queue=queues, var_cost=[10.0], var_current=[n], var_desired=[0], var_pending=[0],
                cfg_kv_threshold=[0.8], cfg_queue_threshold=[5.0], cfg_kv_trigger=[0.1], cfg_queue_trigger=[3.0])


@pytest.mark.parametrize("kvs,queues,up", [([.75, .76], [2, 2], True), ([.5, .5], [3, 3], True),
                                           ([.5, .5], [1, 1], False)])
def test_saturation_scale_up(oracle, kvs, queues, up):
    out = oracle.saturation_v1(_sat_single(kvs, queues))
    assert bool(out["mod_flags"][0] & 1) is up


@pytest.mark.parametrize("kvs,queues,safe", [([.2, .3, .25], [1, 1, 1], True), ([.7, .75], [2, 2], False),
                                             ([.5], [2], False)])
def test_saturation_scale_down_safety(oracle, kvs, queues, safe):
    out = oracle.saturation_v1(_sat_single(kvs, queues))
    assert bool(out["mod_flags"][0] & 2) is safe


# ---- analyzer_test.go:142-192,228-323 aggregation --------------------------------------------------------------------------------------------
def test_saturation_multi_variant_and_saturated_replicas(oracle):
    d = dict(n_models=1, n_variants=2, n_replicas=4, model_variant_off=[0, 2], variant_replica_off=[0, 2, 4],
             rep_kv=[.70, .75, .60, .65], rep_queue=[2, 3, 1, 2], var_cost=[10.0, 10.0], var_current=[2, 2],
             var_desired=[0, 0], var_pending=[0, 0], cfg_kv_threshold=[0.8], cfg_queue_threshold=[5.0],
             cfg_kv_trigger=[0.1], cfg_queue_trigger=[3.0])
    out = oracle.saturation_v1(d)
    assert out["mod_total_replicas"][0] == 4 and out["mod_non_saturated"][0] == 4
    assert out["var_replica_count"].tolist() == [2, 2]
    out = oracle.saturation_v1(_sat_single([.85, .50, .60], [2, 6, 2]))
    assert out["rep_saturated"].tolist() == [1, 1, 0] and out["var_non_saturated"][0] == 1
    assert out["var_max_kv"][0] == .85 and out["var_max_queue"][0] == 6
    empty = _sat_single([], [])
    empty["var_current"] = [0]
    out = oracle.saturation_v1(empty)
    assert out["mod_total_replicas"][0] == 0 and out["mod_flags"][0] == 0


# ---- analyzer_test.go:367-509 CalculateSaturationTargets (exact integers) -----------------------------------------------------------------------
def _targets(oracle, costs, counts, current, desired, kv, queue, pending=None):
    """variants are indexed in ascending-name order: v1-expensive, v2-cheap, v3-medium"""
    V = len(costs)
    off = np.concatenate([[0], np.cumsum(counts)])
    P = int(off[-1])
    d = dict(n_models=1, n_variants=V, n_replicas=P, model_variant_off=[0, V], variant_replica_off=off,
             rep_kv=[kv] * P, rep_queue=[queue] * P, var_cost=costs, var_current=current, var_desired=desired,
             var_pending=pending or [0] * V, cfg_kv_threshold=[0.8], cfg_queue_threshold=[5.0],
             cfg_kv_trigger=[0.1], cfg_queue_trigger=[3.0])
    return oracle.saturation_v1(d)["var_target"].tolist()


def test_saturation_targets(oracle):
    up = dict(kv=0.75, queue=2)      # avg spare KV 0.05 < 0.1 -> ShouldScaleUp
    down = dict(kv=0.2, queue=1)     # ample headroom -> ScaleDownSafe, no scale-up
    assert _targets(oracle, [20., 5., 15.], [2, 2, 2], [2, 2, 2], [0, 0, 0], **up) == [2, 3, 2]      # cheapest +1
    assert _targets(oracle, [20., 5., 15.], [2, 2, 2], [2, 2, 2], [0, 0, 0], **down) == [1, 2, 2]    # most expensive -1
    assert _targets(oracle, [20., 5.], [2, 2], [2, 2], [4, 0], **up) == [4, 2]                       # desired != current blocks
    assert _targets(oracle, [20., 5.], [2, 2], [3, 2], [0, 0], **up) == [3, 2]                       # metrics != current blocks
    assert _targets(oracle, [20., 5., 15.], [2, 2, 2], [2, 2, 2], [0, 0, 0], pending=[0, 1, 0], **up) == [2, 2, 3]
    assert _targets(oracle, [5., 5.], [2, 2], [2, 2], [0, 0], **up) == [3, 2]                        # tie -> first name
    assert _targets(oracle, [5., 5.], [2, 2], [2, 2], [0, 0], **down) == [2, 1]                      # tie -> last name
    assert _targets(oracle, [20., 5.], [1, 1], [1, 1], [0, 0], **down) == [1, 1]                     # never below 1


# ---- internal/engines/pipeline/greedy_saturation_algorithm_test.go:54-270, default_limiter_test.go:152-335 ----------------------------------------
def test_limiter_reference_vectors(oracle):
    g = oracle.limit(dict(n_types=1, acc_type=[0, 0, 0], current=[1, 1, 1], target=[2, 2, 2],
                          gpus_per_replica=[2, 2, 2], spare=[0.3, 0.05, 0.5], cost=[10., 10., 10.], type_limit=[12]))
    assert g["gpus_allocated"].tolist() == [2, 2, 2] and not g["was_limited"].any()
    g = oracle.limit(dict(n_types=1, acc_type=[0, 0, 0], current=[1, 1, 1], target=[2, 2, 2],
                          gpus_per_replica=[2, 2, 2], spare=[0.3, 0.05, 0.5], cost=[10., 10., 10.], type_limit=[10]))
    assert g["target"].tolist() == [2, 2, 1] and g["was_limited"].tolist() == [0, 0, 1]   # order .05, .3, .5
    g = oracle.limit(dict(n_types=1, acc_type=[0], current=[1], target=[3], gpus_per_replica=[2], spare=[0.1],
                          cost=[5.], type_limit=[5]))
    assert g["gpus_allocated"].tolist() == [2] and g["target"].tolist() == [2] and g["was_limited"].tolist() == [1]
    g = oracle.limit(dict(n_types=1, acc_type=[0, 0], current=[1, 1], target=[2, 2], gpus_per_replica=[2, 2],
                          spare=[0.2, 0.2], cost=[20., 5.], type_limit=[6]))
    assert g["target"].tolist() == [1, 2]
    g = oracle.limit(dict(n_types=1, acc_type=[0], current=[1], target=[3], gpus_per_replica=[0], spare=[0.1],
                          cost=[5.], type_limit=[10]))
    assert g["target"].tolist() == [3] and g["gpus_allocated"].tolist() == [2]
    # per-type pools are independent (default_limiter_test.go:278-335)
    g = oracle.limit(dict(n_types=2, acc_type=[0, 1], current=[1, 1], target=[3, 3], gpus_per_replica=[1, 1],
                          spare=[0.1, 0.2], cost=[5., 5.], type_limit=[1 + 2, 1 + 0]))
    assert g["target"].tolist() == [3, 1] and g["was_limited"].tolist() == [0, 1]
    # no scale-up -> untouched; AcceleratorName "" -> nothing allocated
    g = oracle.limit(dict(n_types=1, acc_type=[0, -1], current=[2, 1], target=[1, 2], gpus_per_replica=[1, 1],
                          spare=[0.1, 0.1], cost=[5., 5.], type_limit=[10]))
    assert g["target"].tolist() == [1, 1] and g["was_limited"].tolist() == [0, 1]


# ---- internal/engines/pipeline/type_inventory_test.go:227-300,387-432 (typeAllocator.TryAllocate sequences) -------------------------------------------
TYPE_ALLOCATOR_CASES = [
    # limits by type, decisions (type, current, target, gpus/replica, spare), expected gpus allocated
    ({"A100": 8, "H100": 16}, [("H100", 0, 4, 1, .1), ("A100", 0, 6, 1, .2), ("H100", 0, 20, 1, .3)], [4, 6, 12]),   # :387-432
    ({"H100": 8}, [("H100", 6, 6, 1, .5), ("H100", 0, 4, 1, .1)], [0, 2]),                                          # :268-298 partial
    ({"A100": 8, "H100": 16}, [("H100", 4, 4, 1, .9), ("A100", 4, 4, 1, .9), ("H100", 0, 4, 1, .1)], [0, 0, 4]),    # :227-266 available = limit - used
    ({"H100": 8}, [("H100", 12, 12, 1, .9), ("H100", 0, 1, 1, .1)], [0, 0]),                                        # :156-177 usage above the limit: available 0
]


def limiter_case(limits, decisions):
    types = sorted(limits)
    return dict(n_types=len(types), acc_type=[types.index(t) for t, *_ in decisions], current=[d[1] for d in decisions],
                target=[d[2] for d in decisions], gpus_per_replica=[d[3] for d in decisions], spare=[d[4] for d in decisions],
                cost=[1.0] * len(decisions), type_limit=[limits[t] for t in types])


@pytest.mark.parametrize("limits,decisions,want", TYPE_ALLOCATOR_CASES)
def test_type_allocator_sequences(oracle, limits, decisions, want):
    g = oracle.limit(limiter_case(limits, decisions))
    assert g["gpus_allocated"].tolist() == want
