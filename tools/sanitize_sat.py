"""V1 saturation kernel under compute-sanitizer: every group size, targets-only and detail, on the ragged batch of
tests/test_gpu_parity.py (staged groups, all three fall-back reasons, partial last group) and on a regular batch;
results are checked against the oracle as well."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
pkg = importlib.import_module("llm-d-workload-variant-autoscaler_b200")
from tests import oracle_lib
from test_gpu_parity import _ragged_groups_batch
orc = oracle_lib.load()
with pkg.Engine(0) as e:
    for G in (2, 1, 4):
        os.environ["WVA_SAT_GROUP"] = str(G)
        for d in (_ragged_groups_batch(pkg, 700 + G), pkg.synth.saturation_batch(1500, 32, stream=11)):
            o = orc.saturation_v1(d)
            g = e.saturation_v1(d)
            assert np.array_equal(g["var_target"], o["var_target"]) and np.array_equal(g["mod_flags"], o["mod_flags"])
            assert np.array_equal(g["rep_saturated"], o["rep_saturated"]) and np.array_equal(g["partials"], o["partials"])
            e.saturation_upload(d); e.saturation_run(detail=False)
            r = e.saturation_fetch(detail=False)
            assert np.array_equal(r["var_target"], o["var_target"]) and np.array_equal(r["partials"], o["partials"])
print("sanitize_sat done")
