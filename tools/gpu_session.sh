#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -k "greedy or reference_scenarios or golden" > gpurun_out/s20_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s20_pytest.log
WVA_SIZER_DEBUG=1 timeout 300 python tools/perf_greedy.py > gpurun_out/s20_greedy.json 2> gpurun_out/s20_greedy.err
tail -3 gpurun_out/s20_pytest.log
python -c "
import json
d=json.load(open('gpurun_out/s20_greedy.json'))
for k,v in d.items():
    if isinstance(v,dict): print(k, v['sweep']['ms'], v['queue']['ms'], v['sweep']['events'], v['sweep']['same'], v['queue']['same'])
"; grep "greedy sweep" gpurun_out/s20_greedy.err | sed -n '1p;7p;13p;19p'
