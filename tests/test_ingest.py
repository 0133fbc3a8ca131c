"""Collector -> SoA ingest + streaming reconcile graph (wva_ingest_*) against the reference's join restated in
tests/collector_ref.py and the oracle's V1 saturation analysis."""
import numpy as np
import pytest

from tests import collector_ref as cr

pytestmark = pytest.mark.gpu


def _registry(fx):
    """sorted names -> indices; slots of a variant in ascending pod name (the canonical order)"""
    mvo, vso, slot_of, var_names = [0], [0], {}, []
    for m in fx["models"]:
        for va in fx["variants"][m]:
            var_names.append(va)
            for p in fx["pods"][va]:
                slot_of[p] = len(slot_of)
            vso.append(len(slot_of))
        mvo.append(len(var_names))
    return np.array(mvo, np.int32), np.array(vso, np.int32), slot_of, var_names


def _vector_to_slots(vec, slot_of):
    """what the Prometheus response parser does: one hash look-up per sample (pod label, else pod_name)"""
    sl = np.array([slot_of.get(lab.get("pod") or lab.get("pod_name") or "", -1) for lab, _ in vec], np.int32)
    va = np.array([v for _, v in vec], np.float64)
    return sl, va


@pytest.mark.parametrize("n_models,vpm,seed,threads", [(50, 6, 1, 0), (400, 32, 2, 0), (3, 1, 3, 0), (400, 32, 4, 5),
                                                          (50, 6, 5, 3)])
def test_ingest_matches_collector_and_oracle(pkg, engine, oracle, n_models, vpm, seed, threads, monkeypatch):
    """threads > 0: wva_ingest_write goes through its two-pass radix partition over that many host threads
    (csrc/ingest_scatter.hpp) whatever the size of the vector; 0: the library chooses (serial at these sizes)."""
    if threads:
        monkeypatch.setenv("WVA_INGEST_THREADS", str(threads))
    fx = cr.fixture(n_models, vpm, seed=seed)
    mvo, vso, slot_of, var_names = _registry(fx)
    M, V = len(mvo) - 1, len(vso) - 1
    g = np.random.default_rng(seed)
    state = {"var_cost": (25.0 * 1.35 ** (np.arange(V) % 16)) * (1.0 + (np.arange(V) % vpm) // 16),
             "var_desired": np.zeros(V, np.int32), "var_pending": (g.random(V) < 0.05).astype(np.int32)}
    cfg = {"cfg_kv_threshold": np.full(M, 0.8), "cfg_queue_threshold": np.full(M, 5.0), "cfg_kv_trigger": np.full(M, 0.1),
           "cfg_queue_trigger": np.full(M, 3.0)}
    # ---- reference side: the join, then the oracle's analysis on the joined records
    recs = cr.collect(fx["kv"], fx["queue"], fx["pod_to_variant"])
    vidx = {n: i for i, n in enumerate(var_names)}
    cnt = np.zeros(V, np.int64)
    for r in recs:
        cnt[vidx[r["VariantName"]]] += 1
    vro = np.zeros(V + 1, np.int64); np.cumsum(cnt, out=vro[1:])
    kv = np.array([r["KvCacheUsage"] for r in recs], np.float64)
    q = np.array([r["QueueLength"] for r in recs], np.int64)
    state["var_current"] = cnt.astype(np.int32)
    state["var_current"][::17] += 1                                   # some models in transition (metrics != current)
    batch = {"n_models": M, "n_variants": V, "n_replicas": len(recs), "model_variant_off": mvo,
             "variant_replica_off": vro.astype(np.int32), "rep_kv": kv, "rep_queue": q, **state, **cfg}
    o = oracle.saturation_v1(batch)
    # ---- product side: columnar staging + one graph launch
    ing = pkg.Ingest(engine, mvo, vso)
    try:
        for cycle in range(2):                                        # the second cycle replays the same graph
            ing.begin()
            ing.write(pkg._abi.VEC_KV_CACHE_USAGE, *_vector_to_slots(fx["kv"], slot_of))
            ing.write(pkg._abi.VEC_QUEUE_LENGTH, *_vector_to_slots(fx["queue"], slot_of))
            for k, v in {**state, **cfg}.items():
                ing.cols[k][:] = v
            r = ing.commit()
            assert np.array_equal(r["var_replica_count"], cnt), cycle
            for k in ("var_target", "var_non_saturated", "mod_flags", "mod_total_replicas", "partials"):
                assert np.array_equal(r[k], o[k]), (k, cycle)
            for k in ("var_avg_spare_kv", "var_avg_spare_queue"):
                assert np.array_equal(r[k].view(np.uint64), o[k].view(np.uint64)), (k, cycle)
        # a cycle in which nothing reports: every variant drops out of the analysis
        ing.begin()
        r = ing.commit()
        assert (r["var_replica_count"] == 0).all() and (r["mod_total_replicas"] == 0).all()
    finally:
        ing.close()


def test_ingest_argument_checks(pkg, engine):
    with pytest.raises(pkg.WvaError):
        pkg.Ingest(engine, [0, 2], [0, 1])                 # offsets inconsistent
    ing = pkg.Ingest(engine, [0, 1], [0, 2])
    try:
        with pytest.raises(pkg.WvaError):
            ing.write(0, [5], [0.1])                       # slot out of range
        with pytest.raises(pkg.WvaError):
            ing.write(7, [0], [0.1])                       # unknown vector
        ing.write(0, [-1, 0], [0.3, 0.4])                  # slot < 0 is skipped
        assert ing.cols["kv"][0] == 0.4 and ing.cols["has"][0] == 1
    finally:
        ing.close()
