#!/bin/bash
mkdir -p gpurun_out
WVA_SIZER_DEBUG=1 timeout 300 python tools/perf_grid.py 0.1 > gpurun_out/s10_grid.json 2> gpurun_out/s10_grid.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/s10_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s10_bench_under_ncu.log 2>&1
cat gpurun_out/s10_grid.json; cat gpurun_out/s10_grid.err | tail -8; wc -l gpurun_out/s10_launches.csv
