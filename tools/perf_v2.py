"""Time the V2 pipeline kernels on a seeded batch (M models x 32 variants x 1..8 replicas) and report algorithmic GB/s.
usage: python tools/perf_v2.py [M=200000]"""
import importlib, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("llm-d-workload-variant-autoscaler_b200")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
g = np.random.default_rng(0xB2005EED)
V = M * 32
mvo = (np.arange(M + 1) * 32).astype(np.int32)
nr = g.integers(1, 9, V)
vro = np.concatenate([[0], np.cumsum(nr)]).astype(np.int32)
P = int(vro[-1])
cap = g.choice([8000, 16000, 32000, 64000], P).astype(np.int64)
d = dict(n_models=M, n_variants=V, n_replicas=P, model_variant_off=mvo, variant_replica_off=vro, rep_total_kv_tokens=cap,
         rep_tokens_in_use=(cap * g.uniform(0, 1.0, P)).astype(np.int64), rep_queue_length=g.integers(0, 10, P).astype(np.int64),
         rep_avg_input_tokens=g.uniform(10, 2000, P), rep_avg_output_tokens=g.uniform(10, 800, P), rep_prefix_hit_rate=g.uniform(0, 0.9, P),
         rep_k2=np.where(g.random(P) < 0.5, -1, (cap * g.uniform(0.3, 1.1, P)).astype(np.int64)), rep_slice_order=None,
         var_current=nr.astype(np.int32), var_pending=np.zeros(V, np.int32), var_fallback_capacity=np.zeros(V),
         cfg_kv_threshold=np.full(M, 0.8), cfg_scale_up_threshold=np.full(M, 0.85), cfg_scale_down_boundary=np.full(M, 0.7),
         sched_queue_size=g.integers(0, 20, M).astype(np.int64), sched_queue_bytes=g.integers(0, 100000, M).astype(np.int64))
out = {"M": M, "V": V, "P": P}
with pkg.Engine(0) as e:
    ts = []
    for _ in range(3):
        o = e.saturation_v2(d); ts.append(e.timing()["saturation_ms"])
    in_b = P * 56 + V * 20 + M * 44
    out_b = P * 25 + V * 36 + M * 40
    out["saturation_v2_ms"] = min(ts)
    out["saturation_v2_alg_GBs"] = (in_b + out_b) / (min(ts) * 1e-3) / 1e9
    out["saturation_v2_read_only_GBs"] = in_b / (min(ts) * 1e-3) / 1e9
    opt = dict(model_variant_off=mvo, mod_required_capacity=o["mod_required_capacity"], mod_spare_capacity=o["mod_spare_capacity"],
               mod_has_result=None, var_current=d["var_current"], var_cost=g.choice([1.0, 2.5, 5.0, 10.0, 15.0], V),
               var_per_replica_capacity=o["var_per_replica_capacity"])
    ts = []
    for _ in range(3):
        t = e.cost_aware_optimize(opt); ts.append(e.timing()["limit_ms"])
    out["cost_aware_ms"] = min(ts); out["cost_aware_alg_GBs"] = (V * 24 + M * 20) / (min(ts) * 1e-3) / 1e9
    out["scaled_up_models"] = int((o["mod_required_capacity"] > 0).sum()); out["scaled_down_models"] = int(((o["mod_required_capacity"] <= 0) & (o["mod_spare_capacity"] > 0)).sum())
    enf = dict(model_variant_off=mvo, mod_scale_to_zero_enabled=(g.random(M) < 0.3).astype(np.uint8), mod_request_count=np.where(g.random(M) < 0.5, 0.0, 5.0),
               mod_request_error=None, var_cost=opt["var_cost"], var_has_cost=None, var_target=t)
    ts = []
    for _ in range(3):
        e.enforce(enf); ts.append(e.timing()["limit_ms"])
    out["enforce_ms"] = min(ts); out["enforce_alg_GBs"] = (V * 16 + M * 14) / (min(ts) * 1e-3) / 1e9
    # the fused call: analyzer (without the per-replica outputs nobody downstream reads) -> optimizer -> enforcer on the device
    import ctypes as C
    abi = pkg._abi
    ist, ost, keep, _o = abi.make_saturation_v2(d)
    for k in ("rep_k1", "rep_effective", "rep_demand", "rep_saturated", "var_ready", "var_total_capacity", "var_total_demand", "var_utilization",
              "mod_total_supply", "mod_total_demand", "mod_utilization"):
        setattr(ost, k, None)
    tgt, app = np.zeros(V, np.int32), np.zeros(M, np.uint8)
    cost = np.ascontiguousarray(opt["var_cost"], np.float64); z = np.ascontiguousarray(enf["mod_scale_to_zero_enabled"], np.uint8)
    rc_ = np.ascontiguousarray(enf["mod_request_count"], np.float64)
    ts = []
    with pkg.Engine(0) as e2:
        for _ in range(3):
            r = e2.lib.wva_pipeline_v2(e2.ctx, C.byref(ist), cost.ctypes.data, None, z.ctypes.data, rc_.ctypes.data, None, C.byref(ost), tgt.ctypes.data, app.ctypes.data)
            assert r == 0
            ts.append(e2.timing()["saturation_ms"])
    out["pipeline_v2_three_kernels_ms"] = min(ts)
    out["pipeline_v2_alg_GBs"] = (P * 56 + V * 44 + M * 60) / (min(ts) * 1e-3) / 1e9
print(json.dumps(out, indent=1))
