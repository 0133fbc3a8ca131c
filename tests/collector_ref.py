"""TEST INFRASTRUCTURE: a literal restatement of the reference's CollectReplicaMetrics join
(internal/collector/replica_metrics.go:78-403) for the fields the V1 saturation path reads, over MockPromAPI-shaped
vectors (test/utils/unitutils.go:224: a list of samples, each a label dict + value), plus the fixture generator.

    collect(vectors, pod_to_variant) -> list of ReplicaMetrics dicts (PodName, VariantName, KvCacheUsage, QueueLength)

Order: the reference appends in Go-map order (random); here ascending (VariantName, PodName), the canonical order the
product uses.  Nothing in this module is imported by the product."""
from __future__ import annotations

import math

import numpy as np

INT64_MIN = -(1 << 63)


def go_int(x: float) -> int:
    """Go int(float64) on amd64: truncation; NaN / out of range -> 1 << 63 (CVTTSD2SI's integer indefinite)"""
    if math.isnan(x) or not (-9223372036854775808.0 < x < 9223372036854775808.0):
        return INT64_MIN
    return int(x)


def _pod(labels):
    return labels.get("pod") or labels.get("pod_name") or ""


def collect(kv_vector, queue_vector, pod_to_variant):
    pod = {}
    for labels, value in kv_vector:                               # replica_metrics.go:120-146
        name = _pod(labels)
        if not name:
            continue
        d = pod.setdefault(name, {})
        d["kv"] = float(value); d["hasKv"] = True
    for labels, value in queue_vector:                            # :149-175
        name = _pod(labels)
        if not name:
            continue
        d = pod.setdefault(name, {})
        d["queue"] = go_int(float(value)); d["hasQueue"] = True
    out = []
    for name, d in pod.items():                                   # :296-396
        if not d.get("hasKv") and not d.get("hasQueue"):
            continue
        va = pod_to_variant.get(name, "")                         # PodVAMapper.FindVAForPod
        if va == "":
            continue
        out.append({"PodName": name, "VariantName": va, "KvCacheUsage": d.get("kv", 0.0) if d.get("hasKv") else 0.0,
                    "QueueLength": d.get("queue", 0) if d.get("hasQueue") else 0})
    out.sort(key=lambda r: (r["VariantName"], r["PodName"]))
    return out


def fixture(n_models, variants_per_model, seed=7, max_pods=8, missing=0.08, stray=0.02, dup=0.02):
    """A deployment (models -> variants -> pods) and one cycle of Prometheus-shaped vectors with everything the join has
    to cope with: pods that report only one metric or none, samples of unknown pods, samples without a pod label, the
    `pod_name` label spelling, duplicate samples (last wins), NaN / huge queue values."""
    g = np.random.default_rng(seed)
    models = [f"model-{m:05d}" for m in range(n_models)]
    variants, pods, pod_to_variant = {}, {}, {}
    for m in models:
        variants[m] = [f"{m}-va-{v:02d}" for v in range(variants_per_model)]
        for va in variants[m]:
            n = int(g.integers(0, max_pods + 1))
            pods[va] = sorted(f"{va}-pod-{int(x):04x}" for x in g.choice(65536, n, replace=False))
            for p in pods[va]:
                pod_to_variant[p] = va
    kv_vec, q_vec = [], []
    allpods = [p for va in pods.values() for p in va]
    for p in allpods:
        if g.random() < 0.03:
            continue                                                           # a pod that reports nothing this cycle
        r = g.random()
        lab = {"pod": p} if g.random() < 0.8 else {"pod_name": p}
        if r > missing / 2:
            kv = float(g.beta(2.0, 3.0)) if g.random() > 0.1 else float(g.uniform(0.8, 1.0))
            kv_vec.append((dict(lab), kv))
            if g.random() < dup:
                kv_vec.append((dict(lab), float(g.random())))                 # duplicate: the later sample wins
        if r < 1.0 - missing / 2:
            q = float(g.poisson(1.5)) + (float(g.poisson(8.0)) if g.random() < 0.05 else 0.0)
            if g.random() < 0.002:
                q = float("nan")
            elif g.random() < 0.002:
                q = 1e30
            elif g.random() < 0.01:
                q += 0.75                                                      # fractional sample: int() truncates
            q_vec.append((dict(lab), q))
    for _ in range(int(len(allpods) * stray) + 2):
        kv_vec.append(({"pod": f"stray-{int(g.integers(1 << 30)):x}"}, 0.5))  # pod of no deployment: skipped (:323-328)
        q_vec.append(({}, 3.0))                                               # no pod label: skipped (:127-129)
    order = g.permutation(len(kv_vec)); kv_vec = [kv_vec[i] for i in order]
    # keep duplicates in their relative order (the permutation above may reorder them: recompute "last wins" per pod below)
    order = g.permutation(len(q_vec)); q_vec = [q_vec[i] for i in order]
    return {"models": models, "variants": variants, "pods": pods, "pod_to_variant": pod_to_variant, "kv": kv_vec, "queue": q_vec}
