"""Time wva_solve (limited capacity, SolveGreedy on the device) for every saturation policy.
usage: python tools/perf_greedy.py [scale=0.1] [capacity_frac=0.6]"""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

pkg = importlib.import_module("llm-d-workload-variant-autoscaler_b200")
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.6
e = pkg.Engine(0)
d = pkg.synth.baseline_config(3, scale=scale)
unl = dict(d); unl["unlimited"] = True
e.load_system(unl); e.calculate(); e.solve()
sol_un = e.solution()
out = {"S": int(d["n_servers"]), "A": int(d["n_acc"]), "capacity_frac": frac}
for pol in ("None", "PriorityExhaustive", "PriorityRoundRobin", "RoundRobin"):
    for delayed in (False, True):
        lim = pkg.synth.limit_capacity(d, sol_un["type_count"], frac)
        lim["saturation_policy"] = pol
        lim["delayed_best_effort"] = delayed
        e.load_system(lim); e.calculate()
        ts = []
        for _ in range(3):
            e.solve(); ts.append(e.timing()["solve_ms"])
        g = e.solution()
        out[f"{pol}{'+delayed' if delayed else ''}"] = {"ms": [round(t, 3) for t in ts],
                                                        "allocated": int((g["state"] == 1).sum())}
print(json.dumps(out, indent=1))
print("summary_ms " + json.dumps({k: (min(v["ms"]) if isinstance(v, dict) else v) for k, v in out.items()}))
