#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "saturation or ingest or golden" > gpurun_out/s27_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s27_pytest.log
tail -2 gpurun_out/s27_pytest.log
timeout 300 python tools/perf_sat.py 1000000 12 2> gpurun_out/s27_sat.err | cut -c1-230
timeout 600 ncu --set full --import-source on --clock-control none -k regex:saturation_kernel -s 3 -c 1 -f -o gpurun_out/r2_sat python tools/perf_sat.py 1000000 4 > gpurun_out/s27_ncu.log 2>&1
ls -la gpurun_out/r2_sat.ncu-rep
