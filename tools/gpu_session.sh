#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -k "saturation or config4 or greedy or reference_scenarios or golden or sharded" > gpurun_out/s3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest.log
timeout 300 python tools/perf_sizer_table.py 0.1 128 > gpurun_out/s3_table128.json 2> gpurun_out/s3_table.err
timeout 300 python tools/perf_sizer_table.py 0.1 64 > gpurun_out/s3_table64.json 2>> gpurun_out/s3_table.err
timeout 600 python tools/perf_sat.py > gpurun_out/s3_sat.json 2> gpurun_out/s3_sat.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:saturation_kernel -c 1 -o gpurun_out/s3_sat_prof -f python tools/perf_sat.py 200000 2 > gpurun_out/s3_ncu.log 2>&1
timeout 300 python tools/perf_greedy.py > gpurun_out/s3_greedy.json 2> gpurun_out/s3_greedy.err
tail -3 gpurun_out/s3_pytest.log; cat gpurun_out/s3_table128.json gpurun_out/s3_table64.json; cat gpurun_out/s3_sat.json; tail -2 gpurun_out/s3_sat.err; cat gpurun_out/s3_greedy.json; tail -3 gpurun_out/s3_greedy.err
