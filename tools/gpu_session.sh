#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "calculate or lane or pool or grid or golden or config" > gpurun_out/s29_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s29_pytest.log
tail -2 gpurun_out/s29_pytest.log
for sc in 0.125 1.0; do WVA_SIZER_DEBUG=1 timeout 300 python tools/perf_sizer_full.py $sc 2>gpurun_out/s29_sizer_$sc.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['pairs'], {k:(round(min(v['ms']),2), v['same']) for k,v in d.items() if isinstance(v,dict)})"; done
timeout 300 python tools/perf_grid.py 2>/dev/null | tail -3
