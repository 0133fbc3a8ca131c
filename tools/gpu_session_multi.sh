#!/bin/bash
# multi-GPU session: run with gpurun --gpus N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/m_smi.txt 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/m_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 tools/run_multi_gpu.py --servers 20000 --acc 16 --batch 32 --check > gpurun_out/m_run${N}.json 2> gpurun_out/m_run${N}.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/m_bench${N}.json 2> gpurun_out/m_bench${N}.err; echo "bench rc=$?" >> gpurun_out/m_bench${N}.err
tail -5 gpurun_out/m_pytest.log; cat gpurun_out/m_run${N}.json; tail -3 gpurun_out/m_run${N}.err; tail -c 400 gpurun_out/m_bench${N}.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/m_bench${N}.json'))
    for k in ('value','ms_per_step','solver_wall_ms','saturation','roofline_hbm','e2e'):
        print(k, json.dumps(d.get(k))[:600])
except Exception as e:
    print('no bench line', e)
PY
