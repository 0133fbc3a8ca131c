#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/m_bench${N}.json 2> gpurun_out/m_bench${N}.err; echo "bench rc=$?" >> gpurun_out/m_bench${N}.err
tail -c 300 gpurun_out/m_bench${N}.err; wc -l gpurun_out/m_bench${N}.json; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/m_bench${N}.json'))
    for k in ('value','ms_per_step','solver_wall_ms','saturation','roofline_hbm','e2e'):
        print(k, json.dumps(d.get(k))[:700])
except Exception as e:
    print('no bench line', e)
PY
