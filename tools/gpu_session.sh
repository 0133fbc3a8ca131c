#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -k "saturation or config4 or greedy or reference_scenarios or golden or ingest or pinned or config3 or lane_kernels" > gpurun_out/s5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s5_pytest.log
timeout 600 python tools/perf_sat.py > gpurun_out/s5_sat.json 2> gpurun_out/s5_sat.err
timeout 300 python tools/perf_greedy.py > gpurun_out/s5_greedy.json 2> gpurun_out/s5_greedy.err
WVA_SIZER_DEBUG=1 timeout 600 python tools/perf_sizer_opts.py 1.0 256 2 > gpurun_out/s5_opts.json 2> gpurun_out/s5_opts.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:saturation_kernel -c 1 -o gpurun_out/s5_sat_prof -f python tools/perf_sat.py 200000 2 > gpurun_out/s5_ncu.log 2>&1
tail -3 gpurun_out/s5_pytest.log; cat gpurun_out/s5_sat.json; tail -2 gpurun_out/s5_sat.err; cat gpurun_out/s5_greedy.json; tail -3 gpurun_out/s5_greedy.err; cat gpurun_out/s5_opts.json; cat gpurun_out/s5_opts.err
