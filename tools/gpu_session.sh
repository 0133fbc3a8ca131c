#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -k "calculate or lane_kernels or config3 or saturation or config4 or ingest" > gpurun_out/s7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s7_pytest.log
WVA_SIZER_DEBUG=1 timeout 600 python tools/perf_sizer_full.py 1.0 > gpurun_out/s7_sizer.json 2> gpurun_out/s7_sizer.err
timeout 600 python tools/perf_sat.py > gpurun_out/s7_sat.json 2> gpurun_out/s7_sat.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:sizer_pool_kernel -c 1 -o gpurun_out/s7_pool_prof -f python tools/perf_sizer_full.py 0.05 > gpurun_out/s7_ncu.log 2>&1
tail -3 gpurun_out/s7_pytest.log; cat gpurun_out/s7_sizer.json; cat gpurun_out/s7_sizer.err | tail -12; cat gpurun_out/s7_sat.json
