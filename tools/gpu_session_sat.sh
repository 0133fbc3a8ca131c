#!/bin/bash
# Saturation kernel session: parity tests of the V1 saturation path, then kernel time and instruction count per variant
# (group size x warps per SM; the library must be built with -DWVA_SAT_VARIANTS for the non-default ones).
O=gpurun_out/sat
mkdir -p $O
[ -n "$SKIP_TESTS" ] || timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -k "saturation or ingest or golden or config4 or smoke" > $O/pytest_sat.log 2>&1; echo "pytest rc=$?" >> $O/pytest_sat.log
tail -5 $O/pytest_sat.log
VARIANTS=${VARIANTS:-1,24 2,16 2,20 2,24 4,8 4,12}
for V in $VARIANTS; do
  G=${V%,*}; W=${V#*,}
  WVA_SAT_GROUP=$G WVA_SAT_WARPS=$W timeout 300 python tools/perf_sat.py 1000000 12 > $O/perf_sat_g${G}_w$W.json 2> $O/perf_sat_g${G}_w$W.err; echo "G=$G W=$W $(cat $O/perf_sat_g${G}_w$W.json | head -c 330)"
done
NCU_VARIANTS=${NCU_VARIANTS:-2,24 2,20 4,12}
for V in $NCU_VARIANTS; do
  G=${V%,*}; W=${V#*,}
  WVA_SAT_GROUP=$G WVA_SAT_WARPS=$W timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum --clock-control none -k regex:saturation_kernel -s 3 -c 1 --csv --log-file $O/ncu_inst_g${G}_w$W.csv python tools/perf_sat.py 1000000 4 > /dev/null 2>&1
  echo "G=$G W=$W"; grep -v "^==" $O/ncu_inst_g${G}_w$W.csv | awk -F'","' '{print $(NF-2), $(NF)}' | tr -d '"' | tail -7
done
