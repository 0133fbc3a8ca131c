"""Lane sizer at BASELINE config 3 shape: table placement x queue order x refill policy.
Usage: WVA_SIZER_DEBUG=1 perf_sizer_opts.py [scale=0.1] [N=256]   (live / slots lines go to stderr, in the order of the runs)"""
import importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("llm-d-workload-variant-autoscaler_b200")
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
N = int(sys.argv[2]) if len(sys.argv) > 2 else 256
d = pkg.synth.baseline_config(3, scale=scale) if N == 256 else pkg.synth.queue_system(int(100000 * scale), 32, N, stream=3)
out = {"pairs": int(d["n_servers"] * d["n_acc"]), "N": N, "runs": []}
with pkg.Engine(0) as e:
    e.load_system(d)
    for table in ([int(x) for x in sys.argv[3].split(',')] if len(sys.argv) > 3 else (0, 2)):
        for sort, gang in ((0, 0), (1, 0), (1, 1), (0, 1)):
            e.set_option(4, table); e.set_option(2, sort); e.set_option(3, gang)
            ts = []
            for _ in range(2):
                e.calculate(); ts.append(e.timing()["calculate_ms"])
            sys.stderr.write(f"^ table={table} sort={sort} gang={gang}\n")
            out["runs"].append({"table": table, "sort": sort, "gang": gang, "ms": min(ts)})
    e.set_option(4, 0); e.set_option(2, -1); e.set_option(3, -1)
print(json.dumps(out))
