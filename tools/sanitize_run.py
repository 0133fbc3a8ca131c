"""Small end-to-end pass for compute-sanitizer (memcheck / racecheck / initcheck): every kernel family once."""
import importlib, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("llm-d-workload-variant-autoscaler_b200")
d = pkg.synth.queue_system(24, 6, 32, stream=5)
with pkg.Engine(0) as e:
    for mode in (0, 1, 2, 3, 4, 5):
        e.set_option(1, mode)
        e.load_system(d); e.calculate()
    e.set_option(2, 1); e.set_option(3, 1)            # probe-sorted queue + gang refill
    for mode in (2, 4, 5):
        e.set_option(1, mode)
        e.load_system(pkg.synth.queue_system(40, 6, 32, stream=6)); e.calculate()
    e.set_option(2, -1); e.set_option(3, -1)
    e.set_option(1, 0)
    e.solve(); un = e.solution()
    e.analyze_grid(40)
    lim = pkg.synth.limit_capacity(d, un["type_count"], 0.5); lim["saturation_policy"] = "PriorityRoundRobin"
    e.load_system(lim); e.calculate(); e.solve(); e.solution()
    lim["saturation_policy"] = "None"; lim["delayed_best_effort"] = True
    e.load_system(lim); e.calculate(); e.solve()
    e.saturation_v1(pkg.synth.saturation_batch(50, 7, stream=9))
    e.limit(pkg.synth.limiter_batch(500, 4, stream=9))
    e.mm1k_eval(np.ones(10, np.float32), np.full(10, 2, np.float32), np.full(10, 20, np.int32))
    for pol in ("PriorityExhaustive", "RoundRobin"):
        lim["saturation_policy"] = pol; lim["delayed_best_effort"] = False
        e.load_system(lim); e.calculate(); e.solve(); e.solution()
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from test_pipeline_v2 import random_v2_batch, random_optimizer_batch, random_enforcer_batch
    e.saturation_v2(random_v2_batch(40, 6, max_variants=6, max_replicas=120))       # staged and unstaged models
    big = pkg.synth.queue_system(300, 6, 16, stream=7)                               # greedy with staged head batches
    e.load_system(big); e.calculate(); e.solve(); c = e.candidates()
    blim = pkg.synth.limit_capacity(big, e.solution()["type_count"], 0.4)
    e.load_system(blim); e.set_candidates(c); e.solve(); e.solution()
    e.saturation_v2(random_v2_batch(60, 3)); e.cost_aware_optimize(random_optimizer_batch(60, 4)); e.enforce(random_enforcer_batch(60, 5))
print("sanitize_run done")
