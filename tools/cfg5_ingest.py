"""BASELINE configs[4]: streaming reconcile — 10 k-model Prometheus-shaped metric batches, batch arrival -> decisions,
INCLUDING the SoA packing (the collector's columnar writer), through wva_ingest_* (one CUDA graph per batch) + the
GPU-count limiter.  Usage: cfg5_ingest.py [models=10000] [batches=60]"""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("llm-d-workload-variant-autoscaler_b200")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 60


def measure(e, M, B, shuffled):
    # deployment: every model 32 variants, 1..8 pods each (the synthetic generator of config 4/5); registry = all pods
    d0 = pkg.synth.saturation_batch(M, 32, stream=500)
    V = d0["n_variants"]
    vso = d0["variant_replica_off"].astype(np.int32)           # every pod of the deployment has a slot
    S = int(vso[-1])
    ing = pkg.Ingest(e, d0["model_variant_off"], vso)
    acc_type, gpr = (np.arange(V) % 8).astype(np.int32), np.ones(V, np.int32)
    lat, parts = [], []
    g = np.random.default_rng(5)
    for b in range(B):
        d = pkg.synth.saturation_batch(M, 32, stream=500)       # same deployment (same seed -> same pods) ...
        kvv = np.random.default_rng(1000 + b).beta(2.0, 3.0, S)  # ... fresh metric values every cycle
        qv = np.random.default_rng(2000 + b).poisson(1.5, S).astype(np.float64)
        # two Prometheus vectors keyed by pod: (slot, value) pairs in response order; 2 % of the pods miss a cycle
        rep = np.flatnonzero(g.random(S) > 0.02).astype(np.int32)
        order = g.permutation(rep.size) if shuffled else np.arange(rep.size)
        slots = rep[order]
        kvs, qs = kvv[slots], qv[slots]
        limit = np.full(8, int(d["var_current"].sum() // 8 + 500), np.int32)
        t0 = time.perf_counter()
        ing.begin()
        ing.write(pkg._abi.VEC_KV_CACHE_USAGE, slots, kvs)
        ing.write(pkg._abi.VEC_QUEUE_LENGTH, slots, qs)
        ing.cols["var_cost"][:] = d["var_cost"]; ing.cols["var_current"][:] = d["var_current"]
        ing.cols["var_desired"][:] = d["var_desired"]; ing.cols["var_pending"][:] = d["var_pending"]
        for k in ("cfg_kv_threshold", "cfg_queue_threshold", "cfg_kv_trigger", "cfg_queue_trigger"):
            ing.cols[k][:] = d[k]
        t1 = time.perf_counter()
        r = ing.commit()
        t2 = time.perf_counter()
        gms = e.timing()["saturation_ms"]
        e.limit({"n_types": 8, "acc_type": acc_type, "current": d["var_current"], "target": np.maximum(r["var_target"], 0),
                 "gpus_per_replica": gpr, "spare": r["var_avg_spare_kv"], "cost": d["var_cost"], "type_limit": limit})
        t3 = time.perf_counter()
        lat.append((t3 - t0) * 1e3); parts.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, gms, (t3 - t2) * 1e3))
    ing.close()
    lat = np.array(lat[5:]); parts = np.array(parts[5:])
    return {"p50_ms": float(np.percentile(lat, 50)), "p99_ms": float(np.percentile(lat, 99)),
            "packing_ms_p50": float(np.percentile(parts[:, 0], 50)), "graph_wall_ms_p50": float(np.percentile(parts[:, 1], 50)),
            "graph_device_ms_p50": float(np.percentile(parts[:, 2], 50)), "limiter_ms_p50": float(np.percentile(parts[:, 3], 50)),
            "pods": S, "variants": int(V)}


if __name__ == "__main__":
    with pkg.Engine(0) as e:
        out = {"models_per_batch": M, "batches": B, "response_in_registry_order": measure(e, M, B, False),
               "response_shuffled": measure(e, M, B, True),
               "note": "batch arrival -> decisions: columnar write of two pod-keyed vectors (slot, value) + per-variant state + "
                       "per-model config into the page-locked arena, ONE CUDA graph launch (H2D, pack, V1 saturation analysis + "
                       "targets, D2H), limiter (wva_limit) on the host-visible results"}
    print(json.dumps(out))
