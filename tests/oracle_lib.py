"""ctypes wrapper around oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

The oracle is the CPU restatement of the reference (see oracle/analyzer.hpp).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import importlib
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
abi = importlib.import_module("llm-d-workload-variant-autoscaler_b200._abi")

_f = C.c_float
_fp = C.POINTER(C.c_float)
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        L = lib
        L.oracle_num_threads.restype = C.c_int
        L.oracle_calculate.argtypes = [C.POINTER(abi.System), C.POINTER(abi.Candidates), C.c_int,
                                       C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.oracle_solve.argtypes = [C.POINTER(abi.System), C.POINTER(abi.Candidates), C.POINTER(abi.Solution)]
        L.oracle_type_cost_f32.argtypes = [C.POINTER(abi.System), C.POINTER(abi.Solution), _fp]
        L.oracle_analyze_grid.argtypes = [C.POINTER(abi.System), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_mm1k_eval.argtypes = [C.c_int64] + [C.c_void_p] * 11
        L.oracle_saturation_v1.argtypes = [C.POINTER(abi.SaturationIn), C.POINTER(abi.SaturationOut)]
        L.oracle_limit.argtypes = [C.c_int64, C.c_int] + [C.c_void_p] * 10
        L.oracle_saturation_v2.argtypes = [C.POINTER(abi.SaturationV2In), C.POINTER(abi.SaturationV2Out)]
        L.oracle_cost_aware_optimize.argtypes = [C.c_int64, C.c_int64] + [C.c_void_p] * 8
        L.oracle_enforce.argtypes = [C.c_int64, C.c_int64] + [C.c_void_p] * 8
        L.oracle_enforce_ranked.argtypes = [C.c_int64, C.c_int64] + [C.c_void_p] * 9
        L.oracle_estimate_capacity_from_params.argtypes = [C.c_longlong, C.c_longlong, C.c_double, C.c_double]
        L.oracle_estimate_capacity_from_params.restype = C.c_longlong
        for fn in ("oracle_prefill_time", "oracle_decode_time", "oracle_iteration_time"):
            getattr(L, fn).argtypes = [_f] * 6
            getattr(L, fn).restype = _f
        L.oracle_within_tolerance.argtypes = [_f, _f, _f]
        L.oracle_binary_search_poly.argtypes = [_f] * 6 + [C.c_int, _f, _fp, _ip]
        L.oracle_statedep_solve.argtypes = [C.c_int, _fp, C.c_int, _fp, C.c_int, _fp, _dp]
        L.oracle_mm1k_solve.argtypes = [C.c_int, _f, _f, _fp, _dp]
        L.oracle_queue_analyze.argtypes = [C.c_int, C.c_int] + [_f] * 6 + [_fp, _fp]
        L.oracle_queue_size.argtypes = [C.c_int, C.c_int] + [_f] * 8 + [_fp, _fp, _fp, _ip]
        L.oracle_transition_penalty.argtypes = [C.c_int, C.c_int, _f, C.c_int, C.c_int, C.c_int, _f]
        L.oracle_transition_penalty.restype = _f

    # ---- system level -------------------------------------------------------------------
    def num_threads(self):
        return int(self.lib.oracle_num_threads())

    def calculate(self, sysd, nthreads=0):
        st, keep = abi.make_system(sysd)
        cst, cand = abi.alloc_candidates(st.n_servers, st.n_acc)
        solves, states = C.c_int64(), C.c_int64()
        rc = self.lib.oracle_calculate(C.byref(st), C.byref(cst), nthreads, C.byref(solves), C.byref(states))
        assert rc == 0
        cand["_solves"] = solves.value
        cand["_states"] = states.value
        return cand

    def solve(self, sysd, cand):
        st, keep = abi.make_system(sysd)
        cst, keep2 = abi.candidates_struct(cand)
        sst, sol = abi.alloc_solution(st.n_servers, st.n_types)
        rc = self.lib.oracle_solve(C.byref(st), C.byref(cst), C.byref(sst))
        assert rc == 0
        tc32 = np.zeros(max(st.n_types, 1), dtype=np.float32)
        self.lib.oracle_type_cost_f32(C.byref(st), C.byref(sst), abi.ptr(tc32))
        sol["type_cost_f32"] = tc32[: st.n_types]
        return sol

    def analyze_grid(self, sysd, R, nthreads=0, full=True):
        st, keep = abi.make_system(sysd)
        S, A = st.n_servers, st.n_acc
        n = S * A * R
        out = {}
        if full:
            out["ok"] = np.zeros(n, dtype=np.uint8)
            for k in ("ttft", "itl", "rho", "tput"):
                out[k] = np.zeros(n, dtype=np.float32)
        out["frontier"] = np.zeros(S * A, dtype=np.int32)
        p = lambda k: out[k].ctypes.data if k in out else None
        rc = self.lib.oracle_analyze_grid(C.byref(st), R, p("ok"), p("ttft"), p("itl"), p("rho"), p("tput"),
                                          p("frontier"), nthreads)
        assert rc == 0
        for k in list(out):
            out[k] = out[k].reshape(S, A, R) if k != "frontier" else out[k].reshape(S, A)
        return out

    def mm1k_eval(self, lam, mu, K):
        lam = np.ascontiguousarray(lam, np.float32); mu = np.ascontiguousarray(mu, np.float32)
        K = np.ascontiguousarray(K, np.int32)
        n = lam.size
        out = {"valid": np.zeros(n, np.uint8)}
        names = ("avg_resp", "avg_wait", "avg_serv", "avg_num", "avg_queue", "throughput", "rho")
        for k in names:
            out[k] = np.zeros(n, np.float32)
        self.lib.oracle_mm1k_eval(n, lam.ctypes.data, mu.ctypes.data, K.ctypes.data, out["valid"].ctypes.data,
                                  *[out[k].ctypes.data for k in names])
        return out

    def saturation_v1(self, d):
        ist, keep = abi.make_saturation_in(d)
        ost, out = abi.alloc_saturation_out(ist.n_models, ist.n_variants, ist.n_replicas)
        rc = self.lib.oracle_saturation_v1(C.byref(ist), C.byref(ost))
        assert rc == 0
        return out

    def limit(self, d):
        D = len(d["current"])
        arr = {k: np.ascontiguousarray(d[k], dt) for k, dt in
               (("acc_type", np.int32), ("current", np.int32), ("target", np.int32),
                ("gpus_per_replica", np.int32), ("spare", np.float64), ("cost", np.float64),
                ("type_limit", np.int32))}
        out = {"target": np.zeros(max(D, 1), np.int32), "gpus_allocated": np.zeros(max(D, 1), np.int32),
               "was_limited": np.zeros(max(D, 1), np.uint8)}
        self.lib.oracle_limit(D, int(d["n_types"]), arr["acc_type"].ctypes.data, arr["current"].ctypes.data,
                              arr["target"].ctypes.data, arr["gpus_per_replica"].ctypes.data,
                              arr["spare"].ctypes.data, arr["cost"].ctypes.data, arr["type_limit"].ctypes.data,
                              out["target"].ctypes.data, out["gpus_allocated"].ctypes.data,
                              out["was_limited"].ctypes.data)
        return {k: v[:D] for k, v in out.items()}

    # ---- V2 pipeline -----------------------------------------------------------------------------
    def saturation_v2(self, d):
        ist, ost, keep, out = abi.make_saturation_v2(d)
        assert self.lib.oracle_saturation_v2(C.byref(ist), C.byref(ost)) == 0
        return out

    def cost_aware_optimize(self, d):
        g = lambda k, dt: None if d.get(k) is None else np.ascontiguousarray(d[k], dt).reshape(-1)
        mvo, rq, sp = g("model_variant_off", np.int32), g("mod_required_capacity", np.float64), g("mod_spare_capacity", np.float64)
        hr, cu, co, ca = g("mod_has_result", np.uint8), g("var_current", np.int32), g("var_cost", np.float64), g("var_per_replica_capacity", np.float64)
        tgt = np.zeros(max(len(cu), 1), np.int32)
        p = lambda a: None if a is None or a.size == 0 else a.ctypes.data
        self.lib.oracle_cost_aware_optimize(len(rq), len(cu), p(mvo), p(rq), p(sp), p(hr), p(cu), p(co), p(ca), tgt.ctypes.data)
        return tgt[:len(cu)]

    def enforce(self, d):
        g = lambda k, dt: None if d.get(k) is None else np.ascontiguousarray(d[k], dt).reshape(-1)
        mvo, z, rc, re = g("model_variant_off", np.int32), g("mod_scale_to_zero_enabled", np.uint8), g("mod_request_count", np.float64), g("mod_request_error", np.uint8)
        co, hc = g("var_cost", np.float64), g("var_has_cost", np.uint8)
        tgt = np.ascontiguousarray(d["var_target"], np.int32).reshape(-1).copy()
        V = tgt.size
        if V == 0:
            tgt = np.zeros(1, np.int32)
        app = np.zeros(max(len(rc), 1), np.uint8)
        p = lambda a: None if a is None or a.size == 0 else a.ctypes.data
        self.lib.oracle_enforce(len(rc), V, p(mvo), p(z), p(rc), p(re), p(co), p(hc), tgt.ctypes.data, app.ctypes.data)
        return tgt[:V], app[:len(rc)]

    def pipeline_v2(self, d, var_cost, s2z, request_count, request_error=None, var_name_rank=None):
        """the composition wva_pipeline_v2 must equal: analyzer -> optimizer (every model has a result) -> enforcer"""
        out = self.saturation_v2(d)
        tgt = self.cost_aware_optimize(dict(model_variant_off=d["model_variant_off"], mod_required_capacity=out["mod_required_capacity"],
                                            mod_spare_capacity=out["mod_spare_capacity"], var_current=d["var_current"], var_cost=var_cost,
                                            var_per_replica_capacity=out["var_per_replica_capacity"])).copy()
        g = lambda a, dt: None if a is None else np.ascontiguousarray(a, dt).reshape(-1)
        mvo, z, rc, re, co, rk = g(d["model_variant_off"], np.int32), g(s2z, np.uint8), g(request_count, np.float64), g(request_error, np.uint8), g(var_cost, np.float64), g(var_name_rank, np.int32)
        V = tgt.size
        t = tgt if V else np.zeros(1, np.int32)
        app = np.zeros(max(len(rc), 1), np.uint8)
        p = lambda a: None if a is None or a.size == 0 else a.ctypes.data
        self.lib.oracle_enforce_ranked(len(rc), V, p(mvo), p(z), p(rc), p(re), p(co), None, p(rk), t.ctypes.data, app.ctypes.data)
        return out, t[:V], app[:len(rc)]

    def estimate_capacity_from_params(self, max_batched_tokens, max_num_seqs, avg_in, avg_out):
        return int(self.lib.oracle_estimate_capacity_from_params(int(max_batched_tokens), int(max_num_seqs), float(avg_in), float(avg_out)))

    # ---- KAT helpers ----------------------------------------------------------------------
    def prefill_time(self, a, b, g, i, o, n):
        return float(self.lib.oracle_prefill_time(a, b, g, i, o, n))

    def decode_time(self, a, b, g, i, o, n):
        return float(self.lib.oracle_decode_time(a, b, g, i, o, n))

    def within_tolerance(self, x, v, tol):
        return bool(self.lib.oracle_within_tolerance(x, v, tol))

    def binary_search_poly(self, xmin, xmax, target, c2=0.0, c1=0.0, c0=0.0, fail_at=None):
        x = C.c_float(); ind = C.c_int()
        rc = self.lib.oracle_binary_search_poly(xmin, xmax, target, c2, c1, c0, 0 if fail_at is None else 1,
                                                0.0 if fail_at is None else fail_at, C.byref(x), C.byref(ind))
        return rc, x.value, ind.value

    STAT = ("valid", "lambda", "rho", "avgRespTime", "avgWaitTime", "avgServTime", "avgNumInSystem",
            "avgQueueLength", "throughput", "avgNumInServers")

    def statedep_solve(self, K, serv_rate, lambdas):
        sr = np.ascontiguousarray(serv_rate, np.float32); lam = np.ascontiguousarray(lambdas, np.float32)
        stats = np.zeros((lam.size, 10), np.float32); p = np.zeros(K + 1, np.float64)
        rc = self.lib.oracle_statedep_solve(K, abi.ptr(sr), sr.size, abi.ptr(lam), lam.size,
                                            stats.ctypes.data_as(_fp), abi.ptr(p))
        rows = [dict(zip(self.STAT, map(float, r))) for r in stats]
        return rc, rows, p

    def mm1k_solve(self, K, lam, mu):
        stats = np.zeros(9, np.float32); p = np.zeros(K + 1, np.float64)
        self.lib.oracle_mm1k_solve(K, lam, mu, abi.ptr(stats), abi.ptr(p))
        return dict(zip(self.STAT[:9], map(float, stats))), p

    METRICS = ("Throughput", "AvgRespTime", "AvgWaitTime", "AvgNumInServ", "AvgPrefillTime", "AvgTokenTime",
               "AvgTTFT", "MaxRate", "Rho")

    def queue_analyze(self, max_batch, max_queue, a, b, g, i, o, rate):
        m = np.zeros(9, np.float32); rng = np.zeros(2, np.float32)
        rc = self.lib.oracle_queue_analyze(max_batch, max_queue, a, b, g, i, o, rate, abi.ptr(m), abi.ptr(rng))
        return rc, dict(zip(self.METRICS, map(float, m))), (float(rng[0]), float(rng[1]))

    def queue_size(self, max_batch, max_queue, a, b, g, i, o, ttft, itl, tps):
        rates = np.zeros(3, np.float32); m = np.zeros(9, np.float32); ach = np.zeros(3, np.float32)
        ns = C.c_int()
        rc = self.lib.oracle_queue_size(max_batch, max_queue, a, b, g, i, o, ttft, itl, tps, abi.ptr(rates),
                                        abi.ptr(m), abi.ptr(ach), C.byref(ns))
        return rc, rates, dict(zip(self.METRICS, map(float, m))), ach, ns.value

    def transition_penalty(self, cur_acc, cur_rep, cur_cost, b_state, b_acc, b_rep, b_cost):
        return float(self.lib.oracle_transition_penalty(cur_acc, cur_rep, cur_cost, b_state, b_acc, b_rep, b_cost))


def build():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


_cached = None


def load():
    global _cached
    if _cached is None:
        so = os.path.join(ROOT, "oracle", "liboracle.so")
        srcs = [os.path.join(ROOT, "oracle", f) for f in os.listdir(os.path.join(ROOT, "oracle"))
                if f.endswith((".hpp", ".cpp"))] + [os.path.join(ROOT, "include", "wva_b200.h")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            build()
        _cached = Oracle(C.CDLL(so))
    return _cached
