"""Committed fixtures (tests/golden/): the reference's own known-answer vectors (reference_kat.json, transcribed with
file:line) and frozen oracle outputs on seeded inputs (oracle_outputs.npz, tests/golden/make_golden.py).

CPU: the oracle reproduces both files.  GPU: the device reproduces both files through the C-ABI — without the oracle in
the loop, so an edit that moved oracle and generators together would still be caught."""
import importlib
import json
import os

import numpy as np
import pytest

from tests import ref_scenarios
from tests.test_oracle_kat import limiter_case
from tests.test_pipeline_v2 import random_enforcer_batch, random_optimizer_batch, random_v2_batch

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KAT = json.load(open(os.path.join(HERE, "reference_kat.json")))
GOLD = np.load(os.path.join(HERE, "oracle_outputs.npz"))
FLOATS = (np.float32, np.float64)


def same(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    if a.dtype != b.dtype or a.shape != b.shape:
        return False
    if a.dtype in (np.float32, np.float64):
        return np.array_equal(a.view(np.uint32 if a.dtype == np.float32 else np.uint64), b.view(np.uint32 if a.dtype == np.float32 else np.uint64))
    return np.array_equal(a, b)


def check(prefix, got, skip=()):
    n = 0
    for k in GOLD.files:
        if k.startswith(prefix + "."):
            name = k[len(prefix) + 1:]
            if name in skip or name not in got:
                continue
            assert same(got[name], GOLD[k]), k
            n += 1
    assert n > 0, prefix


def run_kat(backend, pkg):
    """backend: object with cost_aware_optimize / enforce / saturation_v1 / saturation_v2 / limit (oracle or Engine)"""
    for c in KAT["cost_aware"]:
        t = backend.cost_aware_optimize(dict(model_variant_off=[0, len(c["current"])], mod_required_capacity=[c["required"]],
                                             mod_spare_capacity=[c["spare"]], var_current=c["current"], var_cost=c["cost"],
                                             var_per_replica_capacity=c["capacity"]))
        assert t.tolist() == c["want"], c["src"]
    for c in KAT["enforcer"]:
        t, a = backend.enforce(dict(model_variant_off=[0, len(c["target"])], mod_scale_to_zero_enabled=[int(c["s2z"])],
                                    mod_request_count=[c["requests"]], mod_request_error=[int(c["error"])], var_cost=c["cost"],
                                    var_has_cost=c["has_cost"], var_target=c["target"]))
        assert t.tolist() == c["want"] and bool(a[0]) == c["applied"], c["src"]
    for c in KAT["saturation_targets"]:
        V = len(c["cost"])
        off = np.concatenate([[0], np.cumsum(c["replicas"])])
        P = int(off[-1])
        out = backend.saturation_v1(dict(n_models=1, n_variants=V, n_replicas=P, model_variant_off=[0, V], variant_replica_off=off,
                                         rep_kv=[c["kv"]] * P, rep_queue=[c["queue"]] * P, var_cost=c["cost"], var_current=c["current"],
                                         var_desired=c["desired"], var_pending=[0] * V, cfg_kv_threshold=[0.8], cfg_queue_threshold=[5.0],
                                         cfg_kv_trigger=[0.1], cfg_queue_trigger=[3.0]))
        assert out["var_target"].tolist() == c["want"], c["src"]
    for c in KAT["limiter"]:
        g = backend.limit(limiter_case(c["limits"], [tuple(x) for x in c["decisions"]]))
        assert g["gpus_allocated"].tolist() == c["want_gpus"], c["src"]
    for c in KAT["median"]:                       # median of the effective capacities of one variant (k2 observed = value)
        n = len(c["values"])
        d = dict(n_models=1, n_variants=1, n_replicas=n, model_variant_off=[0, 1], variant_replica_off=[0, n],
                 rep_total_kv_tokens=[10**6] * n, rep_tokens_in_use=[0] * n, rep_queue_length=[0] * n, rep_avg_input_tokens=[0.0] * n,
                 rep_avg_output_tokens=[0.0] * n, rep_prefix_hit_rate=[0.0] * n, rep_k2=c["values"], rep_slice_order=None,
                 var_current=[n], var_pending=[0], var_fallback_capacity=[0.0], cfg_kv_threshold=[0.8], cfg_scale_up_threshold=[0.85],
                 cfg_scale_down_boundary=[0.7], sched_queue_size=None, sched_queue_bytes=None)
        assert backend.saturation_v2(d)["var_per_replica_capacity"][0] == float(c["want"]), c["src"]
    for c in KAT["estimate_capacity_from_params"]:
        assert pkg.pipeline.estimate_capacity_from_params({"EffectiveMaxBatchedTokens": c["B"], "MaxNumSeqs": c["S"]}, c["I"], c["O"]) == c["want"], c["src"]


def run_frozen(pkg, calculate, solve, backend):
    for name, d in (("cfg1", pkg.synth.baseline_config(1)), ("mixed", pkg.synth.queue_system(24, 6, 32, stream=5))):
        cand = calculate(d)
        check(f"{name}.cand", cand, skip=("n_solves",))
        sol = solve(d, cand)
        check(f"{name}.sol", sol, skip=("type_cost",))
        if name == "mixed":
            for pol in ("None", "PriorityExhaustive", "PriorityRoundRobin", "RoundRobin"):
                lim = pkg.synth.limit_capacity(d, sol["type_count"], 0.5)
                lim["saturation_policy"] = pol
                check(f"{name}.greedy.{pol}", solve(lim, cand), skip=("type_cost",))
    for sname, (spec, _, _, _) in ref_scenarios.greedy_scenarios().items():
        d, _ = pkg.manager.flatten_spec(spec)
        cand = calculate(d)
        check(f"scenario.{sname}.cand", cand, skip=("n_solves",))
        check(f"scenario.{sname}.sol", solve(d, cand), skip=("type_cost",))
    check("sat_v1", backend.saturation_v1(pkg.synth.saturation_batch(60, 7, stream=9)))
    check("limit", backend.limit(pkg.synth.limiter_batch(800, 5, stream=9, tightness=0.6)))
    check("sat_v2", backend.saturation_v2(random_v2_batch(300, 41)))
    assert same(backend.cost_aware_optimize(random_optimizer_batch(300, 42)), GOLD["cost_aware.target"])
    t, a = backend.enforce(random_enforcer_batch(300, 43))
    assert same(t, GOLD["enforce.target"]) and same(a, GOLD["enforce.applied"])


def test_oracle_reproduces_reference_kat(pkg, oracle):
    run_kat(oracle, pkg)


def test_oracle_reproduces_frozen_outputs(pkg, oracle):
    run_frozen(pkg, oracle.calculate, oracle.solve, oracle)


@pytest.mark.gpu
def test_device_reproduces_reference_kat(pkg, engine):
    run_kat(engine, pkg)


@pytest.mark.gpu
def test_device_reproduces_frozen_outputs(pkg, engine):
    def calculate(d):
        engine.load_system(d); engine.calculate()
        return engine.candidates()

    def solve(d, cand):
        engine.load_system(d); engine.calculate(); engine.solve()
        return engine.solution()
    run_frozen(pkg, calculate, solve, engine)
