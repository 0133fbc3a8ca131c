#!/bin/bash
mkdir -p gpurun_out
for T in 0 1; do for sc in 0.125 1.0; do WVA_POOL_TABLE=$T timeout 300 python tools/perf_sizer_full.py $sc 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('table=$T', d['pairs'], {k:(round(min(v['ms']),2), v['same']) for k,v in d.items() if isinstance(v,dict) and k!='default'})"; done; done
