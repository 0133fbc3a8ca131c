#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gsw_sweep_kernel -s 2 -c 1 -f -o gpurun_out/r2_sweep python tools/perf_greedy.py > gpurun_out/s33_ncu.log 2>&1
tail -2 gpurun_out/s33_ncu.log | cut -c1-300; ls -la gpurun_out/r2_sweep.ncu-rep
