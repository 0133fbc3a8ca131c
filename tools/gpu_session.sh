#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/s8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s8_pytest.log
timeout 600 python tools/cfg5_ingest.py > gpurun_out/s8_cfg5.json 2> gpurun_out/s8_cfg5.err
timeout 900 python bench.py > gpurun_out/s8_bench.json 2> gpurun_out/s8_bench.err; echo "bench rc=$?" >> gpurun_out/s8_bench.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s8_smoke.log 2>&1
tail -3 gpurun_out/s8_pytest.log; cat gpurun_out/s8_cfg5.json; tail -3 gpurun_out/s8_cfg5.err; tail -c 300 gpurun_out/s8_bench.err; cat gpurun_out/s8_smoke.log | tail -2
python - <<PY
import json
d=json.load(open('gpurun_out/s8_bench.json'))
for k in ('value','ms_per_step','solver_wall_ms','greedy','e2e','saturation','roofline_hbm','cpu_baseline','clocks'):
    print(k, json.dumps(d.get(k))[:500])
print('fp64', json.dumps(d['roofline']['fp64'])[:300])
PY
