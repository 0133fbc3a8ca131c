"""Prototype (CPU): lock-step slot efficiency of the lane sizer under different scheduling policies, from the per-solve
chain lengths of real pairs (tests/host_emul emul_trace_pair).  live = sum of chain lengths; slots = 32 x the longest
chain of the warp, per round.  Usage: lockstep_sim.py [servers=600] [N=256]"""
import ctypes as C, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("llm-d-workload-variant-autoscaler_b200")
abi = pkg._abi
S = int(sys.argv[1]) if len(sys.argv) > 1 else 600
N = int(sys.argv[2]) if len(sys.argv) > 2 else 256
d = pkg.synth.queue_system(S, 32, N, stream=3, R=256)
lib = C.CDLL(os.path.join(ROOT, "tests", "host_emul", "libemul.so"))
st, keep = abi.make_system(d)
W = 260
tr = np.zeros((S * 32, W), np.int32)
lib.emul_trace_pair.argtypes = [C.POINTER(abi.System), C.c_void_p, C.c_int]
lib.emul_trace_pair(C.byref(st), tr.ctypes.data, W)
traces = [tr[i, 1:1 + tr[i, 0]].astype(np.int64) for i in range(S * 32) if tr[i, 0] > 0]
print("pairs with solves", len(traces), "solves/pair", np.mean([len(t) for t in traces]), "states/solve", np.mean(np.concatenate(traces)))
live = sum(int(t.sum()) for t in traces)


def sim(order, gang=False, lanes=32, warps=64, skip=0, regroup=0):
    """persistent warps pulling pairs in `order`; returns slots.  skip: drop the first `skip` solves of every pair;
    regroup: per round, lanes of `regroup` warps exchange their solves so that warps hold similar lengths"""
    q = list(order)[::-1]
    nw = warps
    cur = [[None] * lanes for _ in range(nw)]   # (trace, idx)
    slots = 0
    done = False
    groups = [list(range(g, min(g + max(regroup, 1), nw))) for g in range(0, nw, max(regroup, 1))]
    while True:
        any_live = False
        # refill
        for w in range(nw):
            idle = [l for l in range(lanes) if cur[w][l] is None]
            if gang and len(idle) < lanes:
                idle = []
            for l in idle:
                while q:
                    t = traces[q.pop()]
                    if len(t) > skip:
                        cur[w][l] = [t, skip]
                        break
        # one round
        for grp in groups:
            lens = []
            for w in grp:
                for l in range(lanes):
                    c = cur[w][l]
                    if c is not None:
                        lens.append(int(c[0][c[1]]))
            if not lens:
                continue
            any_live = True
            if regroup:
                lens.sort(reverse=True)
                for k in range(0, len(lens), lanes):
                    slots += lanes * lens[k]
            else:
                for w in grp:
                    ls = [int(c[0][c[1]]) for c in cur[w] if c is not None]
                    if ls:
                        slots += lanes * max(ls)
            for w in grp:
                for l in range(lanes):
                    c = cur[w][l]
                    if c is not None:
                        c[1] += 1
                        if c[1] >= len(c[0]):
                            cur[w][l] = None
        if not any_live and not q:
            break
    return slots


nat = list(range(len(traces)))
key = [-int(np.median(t[-4:])) for t in traces]        # stands for the probe: the length the search converges to
srt = sorted(nat, key=lambda i: key[i])
for name, kw in [("natural, individual refill", dict(order=nat)), ("sorted, individual refill", dict(order=srt)),
                 ("sorted, gang refill", dict(order=srt, gang=True)),
                 ("sorted, individual, end points in their own pass", dict(order=srt, skip=2)),
                 ("natural, regroup 8 warps per round", dict(order=nat, regroup=8)),
                 ("sorted, regroup 8 warps per round", dict(order=srt, regroup=8)),
                 ("sorted, regroup 16 warps per round", dict(order=srt, regroup=16))]:
    sl = sim(**kw)
    lv = live if not kw.get("skip") else sum(int(t[kw["skip"]:].sum()) for t in traces)
    print(f"{name:55s} live/slots = {lv / sl:.3f}   slots = {sl / 1e9:.3f} G" + (f"  (+ end-point pass {sum(int(t[:2].sum()) for t in traces) / 1e9:.3f} G live at ~0.95)" if kw.get("skip") else ""))
