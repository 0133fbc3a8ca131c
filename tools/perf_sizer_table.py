"""Lane sizer at BASELINE config 3 shape (N = 256): head table in shared memory (192 lanes / SM) vs global memory
(2 x 256 lanes / SM).  Usage: perf_sizer_table.py [scale=0.1]"""
import importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("llm-d-workload-variant-autoscaler_b200")
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
N = int(sys.argv[2]) if len(sys.argv) > 2 else 256
d = pkg.synth.baseline_config(3, scale=scale) if N == 256 else pkg.synth.queue_system(int(100000 * scale), 32, N, stream=3)
out = {"pairs": int(d["n_servers"] * d["n_acc"]), "N": N}
with pkg.Engine(0) as e:
    e.load_system(d)
    ref = None
    for mode in (0, 2):
        e.set_option(4, mode)
        ts = []
        for _ in range(3):
            e.calculate(); ts.append(e.timing()["calculate_ms"])
        t = e.timing()
        c = e.candidates()
        if ref is None:
            ref = c
        same = all(np.array_equal(np.asarray(c[k]).view(np.uint8), np.asarray(ref[k]).view(np.uint8)) for k in c if k != "n_solves")
        out[f"table_mode_{mode}"] = {"ms": ts, "solves": t["chain_solves"], "states": t["chain_states"], "same_as_mode0": bool(same)}
    e.set_option(4, 0)
print(json.dumps(out))
