"""Host-side mirrors of the two pipeline interfaces the hot path sits behind, above the C-ABI.

* `SaturationAnalyzer` = `interfaces.SaturationAnalyzer` (internal/interfaces/saturation_analyzer.go:246-267), the V1
  percentage analyzer of internal/saturation/analyzer.go;
* `Limiter` = `pipeline.Limiter` (internal/engines/pipeline/limiter_interfaces.go:72-80) as `DefaultLimiter` with
  `TypeInventory` + `GreedyBySaturation` implements it (default_limiter.go:42-113).

Records are plain dicts carrying the reference's Go field names (`PodName`, `KvCacheUsage`, `QueueLength`,
`VariantName`, `AcceleratorName`, `Cost`; `VariantName`/`CurrentReplicas`/`DesiredReplicas`/`PendingReplicas`;
`TargetReplicas`, `GPUsPerReplica`, `SpareCapacity`, `GPUsAllocated`, `WasLimited`, `LimitedBy`, `DecisionSteps`), so a
case from analyzer_test.go / default_limiter_test.go transcribes field by field.  Python twin of
`go/wvab200/wvab200.go`; grouping, ordering and strings only — the arithmetic is `wva_saturation_v1` / `wva_limit`.
"""
from __future__ import annotations

import numpy as np

SAT_SCALE_UP, SAT_SCALE_DOWN_SAFE = 1, 2


def _config(cfg: dict):
    g = lambda k: float(cfg.get(k, 0.0))
    return g("KvCacheThreshold"), g("QueueLengthThreshold"), g("KvSpareTrigger"), g("QueueSpareTrigger")


def scale_up_reason(avg_kv: float, avg_q: float, kv_trigger: float, q_trigger: float) -> str:
    """analyzer.go:199-225 (strings only; the flag itself comes from the device)."""
    kv_t, q_t = avg_kv < kv_trigger, avg_q < q_trigger
    if kv_t and q_t:
        return "both KV spare (%.3f < %.3f) and queue spare (%.1f < %.1f)" % (avg_kv, kv_trigger, avg_q, q_trigger)
    if kv_t:
        return "KV spare Saturation low (%.3f < %.3f)" % (avg_kv, kv_trigger)
    if q_t:
        return "queue spare Saturation low (%.1f < %.1f)" % (avg_q, q_trigger)
    return ""


class SaturationAnalyzer:
    """AnalyzeModelSaturation / CalculateSaturationTargets for one model per call, or `analyze_batch` for every model
    of a reconcile cycle in one launch (what the engine loop at engines/saturation/engine.go:779-795 would hoist)."""

    def __init__(self, engine):
        self.engine = engine

    # -- batching: variants ascending by name inside a model, replica order kept (sums are order dependent) ----
    @staticmethod
    def _pack(models):
        """models: list of (replica_metrics, config, variant_states|None) -> (SoA batch, per-model variant names,
        per-variant metrics lists)."""
        mvo, vro = [0], [0]
        kv, q, cost, cur, des, pen, has = [], [], [], [], [], [], []
        cfg4 = [[], [], [], []]
        names, groups = [], []
        for rm, cfg, states in models:
            by_var = {}
            for r in rm:
                by_var.setdefault(r["VariantName"], []).append(r)
            st = {s["VariantName"]: s for s in (states or [])}
            vnames = sorted(set(by_var) | set(st))        # a state without metrics is still a variant of the model
            for v in vnames:
                ms = by_var.get(v, [])
                kv += [float(r["KvCacheUsage"]) for r in ms]
                q += [int(r["QueueLength"]) for r in ms]
                vro.append(len(kv))
                cost.append(float(ms[0].get("Cost", 0.0)) if ms else 0.0)    # analyzer.go:146-148: first replica's cost
                s = st.get(v)
                cur.append(int(s["CurrentReplicas"]) if s else 0)
                des.append(int(s.get("DesiredReplicas", 0)) if s else 0)
                pen.append(int(s.get("PendingReplicas", 0)) if s else 0)
                has.append(1 if s else 0)
            mvo.append(len(vro) - 1)
            for i, c in enumerate(_config(cfg)):
                cfg4[i].append(c)
            names.append(vnames)
            groups.append(by_var)
        d = dict(n_models=len(models), n_variants=len(vro) - 1, n_replicas=len(kv), model_variant_off=mvo,
                 variant_replica_off=vro, rep_kv=kv, rep_queue=q, var_cost=cost, var_current=cur, var_desired=des,
                 var_pending=pen, var_has_state=has, cfg_kv_threshold=cfg4[0], cfg_queue_threshold=cfg4[1],
                 cfg_kv_trigger=cfg4[2], cfg_queue_trigger=cfg4[3])
        return d, names, groups

    def analyze_batch(self, models):
        """-> list of (ModelSaturationAnalysis dict, targets dict) in input order; one device launch."""
        models = list(models)
        if not models:
            return []
        d, names, groups = self._pack([(m["replicaMetrics"], m["config"], m.get("variantStates")) for m in models])
        out = self.engine.saturation_v1(d)
        res = []
        for mi, m in enumerate(models):
            cfg = _config(m["config"])
            flags = int(out["mod_flags"][mi])
            an = {"ModelID": m.get("modelID", ""), "Namespace": m.get("namespace", ""),
                  "TotalReplicas": int(out["mod_total_replicas"][mi]),
                  "NonSaturatedCount": int(out["mod_non_saturated"][mi]),
                  "AvgSpareKvCapacity": float(out["mod_avg_spare_kv"][mi]),
                  "AvgSpareQueueLength": float(out["mod_avg_spare_queue"][mi]),
                  "ShouldScaleUp": bool(flags & SAT_SCALE_UP), "ScaleUpReason": "",
                  "ScaleDownSafe": bool(flags & SAT_SCALE_DOWN_SAFE), "VariantAnalyses": []}
            if an["ShouldScaleUp"]:
                an["ScaleUpReason"] = scale_up_reason(an["AvgSpareKvCapacity"], an["AvgSpareQueueLength"], cfg[2], cfg[3])
            targets = {}
            v0 = d["model_variant_off"][mi]
            for k, v in enumerate(names[mi]):
                vi = v0 + k
                ms = groups[mi].get(v, [])
                if ms:
                    r0, r1 = d["variant_replica_off"][vi], d["variant_replica_off"][vi + 1]
                    an["VariantAnalyses"].append({
                        "VariantName": v, "AcceleratorName": ms[0].get("AcceleratorName", ""),
                        "Cost": float(d["var_cost"][vi]), "ReplicaCount": int(out["var_replica_count"][vi]),
                        "NonSaturatedCount": int(out["var_non_saturated"][vi]),
                        "MaxKvCacheUsage": float(out["var_max_kv"][vi]), "MaxQueueLength": int(out["var_max_queue"][vi]),
                        "AvgSpareKvCapacity": float(out["var_avg_spare_kv"][vi]),
                        "AvgSpareQueueLength": float(out["var_avg_spare_queue"][vi]),
                        "SaturatedReplicas": [ms[j - r0].get("PodName", "") for j in range(r0, r1) if out["rep_saturated"][j]]})
                if int(out["var_target"][vi]) >= 0:
                    targets[v] = int(out["var_target"][vi])
            res.append((an, targets))
        return res

    def analyze_model_saturation(self, model_id, namespace, replica_metrics, config):
        """AnalyzeModelSaturation (analyzer.go:29-133).  No metrics -> the empty analysis of analyzer.go:39-50."""
        if not replica_metrics:
            return {"ModelID": model_id, "Namespace": namespace, "TotalReplicas": 0, "NonSaturatedCount": 0,
                    "AvgSpareKvCapacity": 0.0, "AvgSpareQueueLength": 0.0, "ShouldScaleUp": False, "ScaleUpReason": "",
                    "ScaleDownSafe": False, "VariantAnalyses": [], "_src": ([], dict(config))}
        an, _ = self.analyze_batch([{"modelID": model_id, "namespace": namespace, "replicaMetrics": replica_metrics,
                                     "config": config}])[0]
        an["_src"] = (list(replica_metrics), dict(config))   # the device computes analysis + targets in one pass
        return an

    def calculate_saturation_targets(self, analysis, variant_states):
        """CalculateSaturationTargets (analyzer.go:296-420): map VariantName -> target replicas."""
        rm, cfg = analysis["_src"]
        if not rm:      # no analyses: every state keeps its current replicas (analyzer.go:303-320 loop over states)
            return {s["VariantName"]: int(s["CurrentReplicas"]) for s in variant_states} if variant_states else {}
        _, targets = self.analyze_batch([{"modelID": analysis["ModelID"], "namespace": analysis["Namespace"],
                                          "replicaMetrics": rm, "config": cfg, "variantStates": variant_states}])[0]
        return targets


class Limiter:
    """DefaultLimiter: `limits` is TypeInventory.limitByType after Refresh (type_inventory.go), a dict accelerator
    type -> GPUs, or a callable returning it.  `limit(decisions)` mutates the decisions in place
    (default_limiter.go:42-113): TargetReplicas, GPUsAllocated, WasLimited, LimitedBy, and one DecisionStep each."""

    def __init__(self, engine, name, limits):
        self.engine, self._name, self._limits = engine, name, limits

    def name(self):
        return self._name

    @staticmethod
    def _reason(d):
        ch = d["TargetReplicas"] - d["CurrentReplicas"]
        if ch <= 0:
            return "no scale-up (target=%d, current=%d)" % (d["TargetReplicas"], d["CurrentReplicas"])
        if d["WasLimited"]:
            return "limited: allocated %d GPUs for +%d replicas" % (d["GPUsAllocated"], ch)
        return "allocated %d GPUs for +%d replicas" % (d["GPUsAllocated"], ch)

    def limit(self, decisions):
        if not decisions:
            return
        lim = self._limits() if callable(self._limits) else self._limits
        types = sorted(lim)
        of = {t: i for i, t in enumerate(types)}
        acc_type = [of[d["AcceleratorName"]] if d.get("AcceleratorName", "") in of and d.get("AcceleratorName", "") != ""
                    else -1 for d in decisions]       # "" or a type without a pool: nothing can be allocated
        out = self.engine.limit({
            "n_types": len(types), "acc_type": acc_type,
            "current": [int(d["CurrentReplicas"]) for d in decisions], "target": [int(d["TargetReplicas"]) for d in decisions],
            "gpus_per_replica": [int(d.get("GPUsPerReplica", 0)) for d in decisions],
            "spare": [float(d.get("SpareCapacity", 0.0)) for d in decisions],
            "cost": [float(d.get("Cost", 0.0)) for d in decisions],
            "type_limit": np.array([int(lim[t]) for t in types], np.int32).reshape(-1)})
        for i, d in enumerate(decisions):
            d["TargetReplicas"] = int(out["target"][i])
            d["GPUsAllocated"] = int(out["gpus_allocated"][i])
            d["WasLimited"] = bool(out["was_limited"][i])
            if d["WasLimited"]:
                d["LimitedBy"] = self._name
            d.setdefault("DecisionSteps", []).append({"Name": self._name, "Action": d.get("Action", ""),
                                                      "TargetReplicas": d["TargetReplicas"], "Reason": self._reason(d),
                                                      "WasConstrained": d["WasLimited"]})
