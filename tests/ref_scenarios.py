"""Systems the reference's own tests build, restated as `config.SystemSpec` JSON-shaped dicts.

Inputs only — the numbers are the ones the reference's tests feed its optimizer (cited per builder); nothing here
computes.  `tests/test_reference_scenarios.py` runs them through the oracle (CPU) and through the C-ABI (GPU) and
checks the assertions the reference's tests make, plus oracle == device bit for bit.
"""
from __future__ import annotations

import copy


def _acc(name, typ, cost, mult=1, mem=40):
    return {"name": name, "type": typ, "multiplicity": mult, "memSize": mem, "cost": cost,
            "power": {"idle": 50, "full": 350, "midPower": 150, "midUtil": 0.4}}


def _perf(model, acc, count, max_batch, at_tokens, alpha, beta, gamma):
    return {"name": model, "acc": acc, "accCount": count, "maxBatchSize": max_batch, "atTokens": at_tokens,
            "serviceParms": {"alpha": alpha, "beta": beta, "gamma": gamma}}


def _server(name, model, cls, rate, in_tok, out_tok, min_rep=1, max_batch=0, keep=False, cur_acc="", cur_rep=0,
            cur_cost=0.0):
    return {"name": name, "class": cls, "model": model, "keepAccelerator": keep, "minNumReplicas": min_rep,
            "maxBatchSize": max_batch,
            "currentAlloc": {"accelerator": cur_acc, "numReplicas": cur_rep, "cost": cur_cost,
                             "load": {"arrivalRate": rate, "avgInTokens": in_tok, "avgOutTokens": out_tok}}}


def _target(model, itl, ttft, tps):
    return {"model": model, "slo-itl": itl, "slo-ttft": ttft, "slo-tps": tps}


def _spec(accs, perf, classes, servers, caps, unlimited=False, policy="None", delayed=False):
    return {"acceleratorData": {"accelerators": accs}, "modelData": {"models": perf},
            "serviceClassData": {"serviceClasses": classes}, "serverData": {"servers": servers},
            "optimizerData": {"optimizer": {"unlimited": unlimited, "delayedBestEffort": delayed,
                                            "saturationPolicy": policy}},
            "capacityData": {"count": [{"type": t, "count": c} for t, c in caps.items()]}}


def single_a100(cur_acc="", cur_rep=0, unlimited=True):
    """pkg/core/system_test.go:42-128 (SetFromSpec), :1186-1246 (Calculate), :1284-1350 (AllocateByType),
    :1392-1458 (GenerateSolution): one A100, one model, class `default` priority 1, one server."""
    return _spec([_acc("A100", "GPU_A100", 1.0)],
                 [_perf("test-model", "A100", 1, 16, 100, 10.0, 2.0, 0.1)],
                 [{"name": "default", "priority": 1, "modelTargets": [_target("test-model", 100, 1000, 50)]}],
                 [_server("test-server", "test-model", "default", 30, 100, 200, 1, 16, cur_acc=cur_acc, cur_rep=cur_rep)],
                 {"GPU_A100": 4}, unlimited=unlimited)


def greedy_base():
    """pkg/solver/greedy_test.go:13-196 setupTestSystemForGreedy: A100 + H100, llama-7b / llama-13b, three classes,
    capacity 4 x GPU_A100 + 2 x GPU_H100, servers server1..server3."""
    accs = [_acc("A100", "GPU_A100", 1.0, 1, 40), _acc("H100", "GPU_H100", 2.0, 1, 80)]
    perf = [_perf("llama-7b", "A100", 1, 16, 100, 10.0, 0.2, 0.01), _perf("llama-7b", "H100", 1, 32, 100, 8.0, 0.15, 0.008),
            _perf("llama-13b", "A100", 2, 8, 150, 15.0, 0.3, 0.01), _perf("llama-13b", "H100", 1, 16, 150, 12.0, 0.25, 0.012)]
    classes = [
        {"name": "high-priority", "priority": 1,
         "modelTargets": [_target("llama-7b", 400, 2000, 15), _target("llama-13b", 500, 2500, 12)]},
        {"name": "medium-priority", "priority": 2,
         "modelTargets": [_target("llama-7b", 450, 2200, 13), _target("llama-13b", 550, 2800, 10)]},
        {"name": "low-priority", "priority": 3, "modelTargets": [_target("llama-7b", 500, 2500, 10)]},
    ]
    servers = [_server("server1", "llama-7b", "high-priority", 30, 100, 200, 1, 512),
               _server("server2", "llama-13b", "medium-priority", 20, 150, 300, 1, 256),
               _server("server3", "llama-7b", "low-priority", 10, 80, 150, 1, 128)]
    return _spec(accs, perf, classes, servers, {"GPU_A100": 4, "GPU_H100": 2})


def _with(base, servers=(), policy="None", delayed=False, caps=None, class_targets=None):
    """AddServerFromSpec replaces a server of the same name (system.go:188-196); AddModelTarget replaces the model's
    target in the class; SetCountFromSpec overwrites the count."""
    s = copy.deepcopy(base)
    by_name = {x["name"]: x for x in s["serverData"]["servers"]}
    for x in servers:
        by_name[x["name"]] = x
    s["serverData"]["servers"] = list(by_name.values())
    s["optimizerData"]["optimizer"].update({"saturationPolicy": policy, "delayedBestEffort": delayed})
    if caps:
        cur = {c["type"]: c["count"] for c in s["capacityData"]["count"]}
        cur.update(caps)
        s["capacityData"]["count"] = [{"type": t, "count": c} for t, c in cur.items()]
    for cls, tgt in (class_targets or {}).items():
        for c in s["serviceClassData"]["serviceClasses"]:
            if c["name"] == cls:
                c["modelTargets"] = [t for t in c["modelTargets"] if t["model"] != tgt["model"]] + [tgt]
    return s


def greedy_scenarios():
    """name -> (spec, servers the reference test inspects, min allocated, max allocated) — greedy_test.go."""
    b = greedy_base()
    hi7 = lambda n, rate=10: _server(n, "llama-7b", "high-priority", rate, 100, 200, 1, 16)
    out = {}
    # :240-294 BasicAllocation — server1 replaced, the llama-7b target of high-priority tightened; only asserts
    # that server1 has candidate allocations
    out["basic"] = (_with(b, [hi7("server1", 30)], "None", False,
                          class_targets={"high-priority": _target("llama-7b", 100, 1000, 50)}), ["server1"], 0, 1)
    # :398-471 PriorityExhaustive, delayed best effort — at least one of server1/server2 allocated
    out["priority_exhaustive"] = (_with(b, [hi7("server1"), hi7("server2")], "PriorityExhaustive", True),
                                  ["server1", "server2"], 1, 2)
    # :473-560 PriorityRoundRobin
    out["priority_round_robin"] = (_with(b, [hi7("server1"), hi7("server2"),
                                             _server("server3", "llama-7b", "medium-priority", 10, 100, 200, 1, 16)],
                                         "PriorityRoundRobin", True), ["server1", "server2", "server3"], 1, 3)
    # :562-649 RoundRobin
    out["round_robin"] = (_with(b, [hi7("server1"), _server("server2", "llama-7b", "medium-priority", 10, 100, 200, 1, 16),
                                    _server("server3", "llama-7b", "low-priority", 10, 100, 200, 1, 16)],
                                "RoundRobin", True), ["server1", "server2", "server3"], 1, 3)
    # :651-718 ResourceExhaustion — 1 + 1 units, five competing servers: some but not all allocated
    out["resource_exhaustion"] = (_with(b, [hi7(f"server{i}", 20) for i in range(1, 6)], "PriorityExhaustive", True,
                                        caps={"GPU_A100": 1, "GPU_H100": 1}),
                                  [f"server{i}" for i in range(1, 6)], 1, 4)
    # :720-814 HighLoadScenario
    out["high_load"] = (_with(b, [_server("server1", "llama-7b", "high-priority", 100, 200, 300, 2, 32),
                                  _server("server2", "llama-7b", "medium-priority", 80, 150, 250, 1, 16),
                                  _server("server3", "llama-13b", "low-priority", 50, 200, 400, 1, 8)],
                              "PriorityExhaustive", True), ["server1", "server2", "server3"], 1, 3)
    # :816-889 MixedModelTypes
    out["mixed_models"] = (_with(b, [_server("llama7b-server", "llama-7b", "high-priority", 40, 100, 200, 1, 16),
                                     _server("llama13b-server", "llama-13b", "high-priority", 30, 150, 300, 1, 8)],
                                 "RoundRobin", True), ["llama7b-server", "llama13b-server"], 1, 2)
    # :891-965 EdgeCases — zero load and very high load
    out["edge_cases"] = (_with(b, [_server("zero-load-server", "llama-7b", "high-priority", 0, 100, 200, 1, 16),
                                   _server("high-load-server", "llama-7b", "medium-priority", 1000, 500, 1000, 3, 64)],
                               "PriorityRoundRobin", True), ["zero-load-server", "high-load-server"], 1, 2)
    return out
