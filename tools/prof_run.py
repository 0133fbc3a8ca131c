"""Small driver for ncu captures: runs the chosen stage of the hot path on BASELINE configs[1]."""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("llm-d-workload-variant-autoscaler_b200")
what = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
S = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
d = pkg.synth.queue_system(S, 16, 128, n_classes=3, stream=2, R=128)
with pkg.Engine(0) as e:
    e.load_system(d)
    if os.environ.get("WVA_MODE"):
        e.set_option(1, int(os.environ["WVA_MODE"]))
    for _ in range(reps):
        if what in ("all", "sizer"):
            e.calculate(); print("calculate", e.timing())
            e.solve()
        if what in ("all", "grid"):
            e.grid_run(128, full=True); print("grid", e.timing())
    if what in ("all", "sat"):
        sd = pkg.synth.saturation_batch(200000, 32)
        e.saturation_upload(sd)
        for _ in range(reps):
            e.saturation_run(False); print("sat", e.timing(), sd["n_replicas"])
