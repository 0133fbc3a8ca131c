#!/bin/bash
# Source-level instruction counts of the saturation kernel (one group size), exported as CSV (the report stays on the box).
G=${1:-2}
O=gpurun_out/sat
mkdir -p $O; R=/tmp/ncu_reps; mkdir -p $R
WVA_SAT_GROUP=$G WVA_SAT_WARPS=${2:-24} timeout 600 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --import-source on --clock-control none -k regex:saturation_kernel -s 3 -c 1 -f -o $R/satsrc python tools/perf_sat.py 1000000 4 > $O/ncu_src.log 2>&1
ncu -i $R/satsrc.ncu-rep --page source --csv --print-source cuda,sass > $O/src_cuda_sass_g$G.csv 2>/dev/null
ncu -i $R/satsrc.ncu-rep --page source --csv --print-source sass > $O/src_sass_g$G.csv 2>/dev/null
ncu -i $R/satsrc.ncu-rep --page raw --csv > $O/src_raw_g$G.csv 2>/dev/null
ls -la $O $R | tail -12
head -c 1500 $O/src_cuda_sass_g$G.csv
