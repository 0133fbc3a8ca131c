#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "saturation or ingest or golden" > gpurun_out/s25_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s25_pytest.log
tail -3 gpurun_out/s25_pytest.log
timeout 300 python tools/perf_sat.py 1000000 12 > gpurun_out/s25_sat.json 2> gpurun_out/s25_sat.err; cat gpurun_out/s25_sat.json
