#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -k "saturation or config4 or config3_slice or pinned or smoke" > gpurun_out/s2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2_pytest.log
WVA_SIZER_DEBUG=1 timeout 300 python tools/perf_sizer_table.py 0.1 > gpurun_out/s2_table.json 2> gpurun_out/s2_table.err
timeout 600 python tools/perf_sat.py > gpurun_out/s2_sat.json 2> gpurun_out/s2_sat.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:saturation_kernel -c 2 -o gpurun_out/s2_sat_prof -f python tools/perf_sat.py 200000 2 > gpurun_out/s2_ncu.log 2>&1
tail -3 gpurun_out/s2_pytest.log; cat gpurun_out/s2_table.json; tail -3 gpurun_out/s2_table.err; cat gpurun_out/s2_sat.json; tail -2 gpurun_out/s2_sat.err
