"""Time wva_solve (limited capacity, SolveGreedy on the device) for every saturation policy, both formulations
(static-order sweep / literal queue).  usage: python tools/perf_greedy.py [servers=100000] [capacity_frac=0.6]"""
import importlib, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

pkg = importlib.import_module("llm-d-workload-variant-autoscaler_b200")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.6
e = pkg.Engine(0)
d = pkg.synth.queue_system(S, 32, 16, stream=3, R=256)      # config-3 loads and SLOs, small N: only the allocator is timed
e.load_system(d); e.calculate(); e.set_optimizer(True); e.solve()
un = e.solution()
cap = np.maximum(1, np.floor(np.asarray(un["type_count"], np.float64) * frac)).astype(np.int32)
e.set_capacity(cap)
out = {"S": S, "A": 32, "capacity_frac": frac}
for pol in ("None", "PriorityExhaustive", "PriorityRoundRobin", "RoundRobin"):
    for delayed in (False, True):
        row = {}
        ref = None
        for mode, name in ((2, "sweep"), (1, "queue")):
            e.set_option(5, mode)
            e.set_optimizer(False, delayed, pol)
            ts = []
            for _ in range(3):
                e.solve(); ts.append(e.timing()["solve_ms"])
            t = e.timing()
            g = e.solution()
            if ref is None:
                ref = g
            same = all(np.array_equal(np.asarray(g[k]).view(np.uint8), np.asarray(ref[k]).view(np.uint8)) for k in g)
            row[name] = {"ms": round(min(ts), 3), "events": t["greedy_events"], "allocated": int((g["state"] == 1).sum()), "same": bool(same)}
        out[f"{pol}{'+delayed' if delayed else ''}"] = row
e.set_option(5, 0)
print(json.dumps(out))
