#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -k "calculate or lane_kernels or config3 or config2 or greedy_at_scale or both_formulations or default_policy" > gpurun_out/s6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s6_pytest.log
WVA_SIZER_DEBUG=1 timeout 600 python tools/perf_sizer_full.py 0.1 > gpurun_out/s6_sizer01.json 2> gpurun_out/s6_sizer01.err
WVA_SIZER_DEBUG=1 timeout 600 python tools/perf_sizer_full.py 1.0 > gpurun_out/s6_sizer.json 2> gpurun_out/s6_sizer.err
tail -3 gpurun_out/s6_pytest.log; cat gpurun_out/s6_sizer01.json; cat gpurun_out/s6_sizer01.err | tail -12; cat gpurun_out/s6_sizer.json; cat gpurun_out/s6_sizer.err | tail -12
