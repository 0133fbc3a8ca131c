#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -k "grid or config2 or smoke" > gpurun_out/s12_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s12_pytest.log
WVA_SIZER_DEBUG=1 timeout 300 python tools/perf_grid.py 1.0 > gpurun_out/s12_grid.json 2> gpurun_out/s12_grid.err
WVA_SIZER_DEBUG=1 timeout 300 python tools/perf_greedy.py > gpurun_out/s12_greedy.json 2> gpurun_out/s12_greedy.err
tail -3 gpurun_out/s12_pytest.log; cat gpurun_out/s12_grid.json; tail -2 gpurun_out/s12_grid.err; grep "greedy sweep" gpurun_out/s12_greedy.err | head -3
