"""System.Calculate at BASELINE configs[2] (or a scaled copy): default path vs forced lane sizer.  Usage: [scale=1.0]"""
import importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("llm-d-workload-variant-autoscaler_b200")
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
d = pkg.synth.baseline_config(3, scale=scale)
out = {"pairs": int(d["n_servers"] * d["n_acc"])}
with pkg.Engine(0) as e:
    e.load_system(d)
    ref = None
    for name, opt in (("default", 0), ("pool", 6), ("lane", 2)):
        e.set_option(1, opt)
        ts = []
        for _ in range(3):
            e.calculate(); ts.append(e.timing()["calculate_ms"])
        t = e.timing()
        sys.stderr.write(f"^ {name}\n")
        c = e.candidates()
        if ref is None:
            ref = c
        same = all(np.array_equal(np.asarray(c[k]).view(np.uint8), np.asarray(ref[k]).view(np.uint8)) for k in c if k != "n_solves")
        out[name] = {"ms": ts, "solves": t["chain_solves"], "states": t["chain_states"], "same": bool(same)}
    e.set_option(1, 0)
print(json.dumps(out))
