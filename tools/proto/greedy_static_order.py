"""Prototype (CPU, numpy/python): Solver.SolveGreedy's allocate() (pkg/solver/greedy.go:107-166) as ONE sweep over a
STATICALLY ordered list of candidate events, checked against the oracle's literal sorted-slice algorithm.

Claim.  allocate() is a priority queue over entries keyed k(e, j) = (priority asc, delta_j desc, value_j desc) of the
entry's CURRENT candidate j, re-inserted BEFORE equal elements.  Event (e, j) — "entry e is tested at candidate j" — can
only happen after (e, j-1) failed, and then happens at queue time tau(e, j) = max_{i <= j} k(e, i) (a key smaller than the
queue head is popped at once).  tau is a property of the entry alone, so all S x A potential events can be sorted ONCE;
the sweep then only needs per-entry "alive" bits and the per-type capacities.  Ties: among equal tau, re-inserted
leaders (key == tau > tau of their predecessor) come first, latest insertion first — resolved at run time from the
stamps of the predecessors' failures, only the ACTIVE ones matter —, then the original entries in canonical order; the
events of one entry with the same tau form a run processed back to back.
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def sortable(x):
    """float32 -> uint32 order-preserving, NaN first, -0 == +0 (Go cmp.Compare)"""
    x = np.asarray(x, np.float32).copy()
    x[x == 0] = 0.0
    b = x.view(np.uint32).astype(np.uint64)
    out = np.where(b & 0x80000000, (~b) & 0xFFFFFFFF, b | 0x80000000)
    out[np.isnan(x)] = 0
    return out.astype(np.uint64)


def static_greedy_none(sysd, cand):
    """policy None, delayed or not (identical without best effort).  Returns (sel_acc[S], unallocated order)."""
    S, A = int(sysd["n_servers"]), int(sysd["n_acc"])
    state = np.asarray(cand["state"]).reshape(S, A)
    value = np.asarray(cand["value"], np.float32).reshape(S, A)
    nrep = np.asarray(cand["num_replicas"]).reshape(S, A).astype(np.int64)
    acc_type = np.asarray(sysd["acc_type"]); mult = np.asarray(sysd["acc_multiplicity"]).astype(np.int64)
    inst = np.maximum(np.asarray(sysd["perf_acc_count"]).reshape(-1, A), 1).astype(np.int64)
    model = np.asarray(sysd["srv_model"]); prio = np.asarray(sysd["srv_priority"]).astype(np.uint64)
    avail = np.asarray(sysd["type_count"]).astype(np.int64).copy()
    ev = []          # (tau_hi, tau_lo, cls, entry, j, leader) ; tau = (prio<<32 | ~sortable(delta), ~sortable(value))
    rec = {}
    for e in range(S):
        idx = [a for a in range(A) if state[e, a] != 0]
        idx.sort(key=lambda a: (0 if np.isnan(value[e, a]) else 1, value[e, a] if not np.isnan(value[e, a]) else 0.0, a))
        n = len(idx)
        if n == 0:
            continue
        vals = value[e, idx]
        tau = None
        for j, a in enumerate(idx):
            d = np.float32(vals[j + 1] - vals[j]) if j + 1 < n else np.float32(np.finfo(np.float32).max)
            khi = (int(prio[e] ^ 0x80000000) << 32) | int((~sortable([d])[0]) & 0xFFFFFFFF)
            klo = int((~sortable([vals[j]])[0]) & 0xFFFFFFFF)
            k = (khi, klo)
            leader = tau is None or k > tau
            if leader:
                tau = k
            cls = 1 if j == 0 else 0                      # originals after re-inserted leaders
            live = model[e] >= 0 and state[e, a] == 1
            t = int(acc_type[a]) if live else -1
            cnt = int(nrep[e, a] * inst[model[e], a] * mult[a]) if live else 0
            rec[(e, j)] = (t, cnt, a, n)
            ev.append((tau[0], tau[1], e, j, leader, cls))
    # static order: tau, then (class of the run's leader, entry, j).  The leader's class is attached to every event of its run.
    runs_cls = {}
    for (thi, tlo, e, j, leader, cls) in ev:
        if leader:
            cur = cls
            runs_cls[(e, j)] = cur
        else:
            runs_cls[(e, j)] = runs_cls[(e, j - 1)]
    ev.sort(key=lambda x: (x[0], x[1], runs_cls[(x[2], x[3])], x[2], x[3]))
    alive = np.ones(S, bool)
    stamp = np.zeros(S, np.int64)
    sel = np.full(S, -1, np.int64)
    unalloc = []
    clock = 0

    def process(e, j):
        nonlocal clock
        if not alive[e]:
            return
        clock += 1
        t, cnt, a, n = rec[(e, j)]
        if t < 0:
            alive[e] = False            # dropped (greedy.go:126-136)
            return
        if avail[t] >= cnt:
            avail[t] -= cnt; sel[e] = a; alive[e] = False
            return
        stamp[e] = clock
        if j == n - 1:
            alive[e] = False; unalloc.append(e)

    i, N = 0, len(ev)
    n_dyn = 0
    while i < N:
        k = i
        while k < N and ev[k][0] == ev[i][0] and ev[k][1] == ev[i][1]:
            k += 1
        grp = ev[i:k]
        # class-0 runs of the group (leaders inserted from lower levels)
        lead0 = [g for g in grp if g[4] and runs_cls[(g[2], g[3])] == 0]
        if len(lead0) >= 2:
            act = [g for g in lead0 if alive[g[2]]]
            if len(act) >= 2:
                n_dyn += 1
            act.sort(key=lambda g: -stamp[g[2]])                  # latest insertion first
            for g in act:
                e, j = g[2], g[3]
                jj = j
                while (e, jj) in rec and jj <= rec[(e, jj)][3] - 1:
                    # events of the run: same entry, same tau
                    if jj > j and not any(x[2] == e and x[3] == jj for x in grp):
                        break
                    process(e, jj)
                    jj += 1
            for g in grp:
                if runs_cls[(g[2], g[3])] == 1:
                    process(g[2], g[3])
        else:
            for g in grp:
                process(g[2], g[3])
        i = k
    return sel, unalloc, n_dyn


def main():
    pkg = importlib.import_module("llm-d-workload-variant-autoscaler_b200")
    from tests import oracle_lib
    orc = oracle_lib.load()
    bad = 0
    cases = []
    for seed, (S, A, N) in enumerate([(300, 8, 16), (96, 6, 16), (500, 12, 8), (200, 4, 16), (1000, 16, 4)]):
        d = pkg.synth.queue_system(S, A, N, stream=300 + seed)
        cases.append(("rand", d))
    d = pkg.synth.queue_system(96, 6, 16, stream=72)
    for k, v in list(d.items()):
        if isinstance(v, np.ndarray) and v.shape[:1] == (96,):
            v[:] = np.concatenate([v[:8]] * 12)
    cases.append(("dups", d))
    d = pkg.synth.queue_system(400, 8, 8, stream=311, zero_load_frac=0.6)
    cases.append(("zeroload", d))
    for name, d in cases:
        c = orc.calculate(d)
        un = dict(d); un["unlimited"] = True
        s0 = orc.solve(un, c)
        for frac in (0.15, 0.3, 0.6, 0.9):
            for delayed in (False, True):
                lim = pkg.synth.limit_capacity(d, s0["type_count"], frac)
                lim["saturation_policy"] = "None"; lim["delayed_best_effort"] = delayed
                o = orc.solve(lim, c)
                sel, un_list, n_dyn = static_greedy_none(lim, c)
                oacc = np.where(o["state"] == 1, o["acc"], -1)
                ok = np.array_equal(sel, oacc)
                bad += not ok
                print(name, d["n_servers"], d["n_acc"], frac, delayed, "OK" if ok else f"MISMATCH {np.flatnonzero(sel != oacc)[:8]}",
                      "dyn groups", n_dyn, "allocated", int((oacc >= 0).sum()))
    print("mismatches:", bad)


if __name__ == "__main__":
    main()
