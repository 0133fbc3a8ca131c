"""V1 saturation kernel at BASELINE config 4 size (1 M models x 32 variants): kernel time, algorithmic GB/s.
Usage: perf_sat.py [models=1000000] [reps=10]"""
import importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("llm-d-workload-variant-autoscaler_b200")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
d = pkg.synth.saturation_batch(M, 32, stream=4)
alg = d["n_replicas"] * 16 + d["n_variants"] * 32 + M * 40
with pkg.Engine(0) as e:
    e.saturation_upload(d)
    out = {}
    for detail in (False, True):
        ks = []
        for _ in range(reps):
            e.saturation_run(detail); ks.append(e.timing()["saturation_ms"])
        out["detail" if detail else "targets_only"] = {"ms_min": min(ks), "ms_median": float(np.median(ks)),
                                                        "alg_gbs": alg / (min(ks) * 1e-3) / 1e9}
    r = e.saturation_fetch(False)
    out.update(models=M, replicas=int(d["n_replicas"]), alg_bytes=int(alg), partials=r["partials"].tolist())
print(json.dumps(out))
