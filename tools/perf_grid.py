"""Replica grid at BASELINE config 3 shape.  Usage: WVA_SIZER_DEBUG=1 perf_grid.py [scale=0.1]"""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("llm-d-workload-variant-autoscaler_b200")
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
d = pkg.synth.baseline_config(3, scale=scale)
with pkg.Engine(0) as e:
    e.load_system(d)
    out = {}
    for full in (False, True):
        ts = []
        for _ in range(3):
            e.grid_run(256, full=full); ts.append(e.timing()["grid_ms"])
        t = e.timing()
        out["full" if full else "frontier"] = {"ms": ts, "solves": t["chain_solves"], "states": t["chain_states"]}
print(json.dumps(out))
