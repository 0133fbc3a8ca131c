#!/bin/bash
# one gpurun call: GPU tests, sizer table-placement probe, the default bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/s1_smi.txt 2>&1
nproc > gpurun_out/s1_host.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/s1_host.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/s1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s1_pytest.log
WVA_SIZER_DEBUG=1 timeout 300 python tools/perf_sizer_table.py 0.1 > gpurun_out/s1_table.json 2> gpurun_out/s1_table.err
timeout 900 python bench.py > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err; echo "bench rc=$?" >> gpurun_out/s1_bench.err
tail -3 gpurun_out/s1_pytest.log; cat gpurun_out/s1_table.json; tail -c 600 gpurun_out/s1_bench.err; head -c 1500 gpurun_out/s1_bench.json
