#!/bin/bash
mkdir -p gpurun_out
for sc in 0.03 0.0625 0.125 0.25 0.5; do timeout 300 python tools/perf_sizer_full.py $sc 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['pairs'], {k:(round(min(v['ms']),2), v['same']) for k,v in d.items() if isinstance(v,dict)})"; done
