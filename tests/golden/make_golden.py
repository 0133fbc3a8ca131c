"""Regenerates tests/golden/*.npz and reference_kat.json's derived part.

The reference is Go and cannot run in this image or on the GPU box (no toolchain), so the golden OUTPUTS here are
produced by the oracle (oracle/, pinned to the reference's own test vectors by tests/test_oracle_kat.py,
tests/test_pipeline_v2.py and tests/test_reference_scenarios.py) on seeded inputs; they freeze the oracle so that a later
edit of oracle/ or of the generators cannot move both sides of a parity test at once.  The hand-transcribed vectors of the
reference's tests live in reference_kat.json (each with its file:line) and are NOT generated.

usage: python tests/golden/make_golden.py        (from the repo root; needs only the CPU)
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("llm-d-workload-variant-autoscaler_b200")
from tests import oracle_lib, ref_scenarios  # noqa: E402
from tests.test_pipeline_v2 import random_enforcer_batch, random_optimizer_batch, random_v2_batch  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def flat(prefix, d):
    return {f"{prefix}.{k}": np.asarray(v) for k, v in d.items() if isinstance(v, np.ndarray)}


def main():
    o = oracle_lib.load()
    out = {}
    # BASELINE configs[0] (the reference's CPU-runnable case) and a 3-class system with zero-load / infeasible servers
    for name, d in (("cfg1", pkg.synth.baseline_config(1)), ("mixed", pkg.synth.queue_system(24, 6, 32, stream=5))):
        cand = o.calculate(d)
        out.update(flat(f"{name}.cand", cand))
        out.update(flat(f"{name}.sol", o.solve(d, cand)))
        if name == "mixed":
            un = o.solve(d, cand)
            for pol in ("None", "PriorityExhaustive", "PriorityRoundRobin", "RoundRobin"):
                lim = pkg.synth.limit_capacity(d, un["type_count"], 0.5)
                lim["saturation_policy"] = pol
                out.update(flat(f"{name}.greedy.{pol}", o.solve(lim, cand)))
    # the reference's greedy test systems
    for sname, (spec, _, _, _) in ref_scenarios.greedy_scenarios().items():
        d, _ = pkg.manager.flatten_spec(spec)
        cand = o.calculate(d)
        out.update(flat(f"scenario.{sname}.cand", cand))
        out.update(flat(f"scenario.{sname}.sol", o.solve(d, cand)))
    out.update(flat("sat_v1", o.saturation_v1(pkg.synth.saturation_batch(60, 7, stream=9))))
    out.update(flat("limit", o.limit(pkg.synth.limiter_batch(800, 5, stream=9, tightness=0.6))))
    out.update(flat("sat_v2", o.saturation_v2(random_v2_batch(300, 41))))
    out["cost_aware.target"] = o.cost_aware_optimize(random_optimizer_batch(300, 42))
    t, a = o.enforce(random_enforcer_batch(300, 43))
    out["enforce.target"], out["enforce.applied"] = t, a
    np.savez_compressed(os.path.join(HERE, "oracle_outputs.npz"), **out)
    print(f"wrote {len(out)} arrays, {sum(v.nbytes for v in out.values()) / 1e3:.0f} kB uncompressed")


if __name__ == "__main__":
    main()
