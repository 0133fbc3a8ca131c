#!/bin/bash
# Final-build evidence of a round: full GPU suite, smoke, the bench line (both arms), the ncu launch list of the bench
# command, and `ncu --set full` summaries of the dominant kernels (text summaries only: the reports are too big to travel).
TAG=${1:-final}
O=gpurun_out/$TAG
mkdir -p $O
git_rev=$(cat .git_rev 2>/dev/null)
echo "build: $git_rev" > $O/build.txt
timeout 1800 python -m pytest tests -q -m gpu -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 900 python bench.py > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "bench rc=$?" >> $O/bench_1gpu.err
timeout 900 python bench.py --impl reference > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/launches_bench.out 2>&1
R=/tmp/ncu_reps; mkdir -p $R
timeout 600 ncu --set full --import-source on --clock-control none -k regex:sizer_pool_kernel -s 1 -c 1 -f -o $R/pool python tools/perf_sizer_full.py 0.125 > $O/ncu_pool.log 2>&1
python tools/ncu_summary.py $R/pool.ncu-rep > $O/ncu_sizer_pool.txt 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:grid_ -s 2 -c 3 -f -o $R/grid python tools/perf_grid.py 0.1 > $O/ncu_grid.log 2>&1
python tools/ncu_summary.py $R/grid.ncu-rep > $O/ncu_grid.txt 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:saturation_kernel -s 3 -c 1 -f -o $R/sat python tools/perf_sat.py 1000000 4 > $O/ncu_sat.log 2>&1
python tools/ncu_summary.py $R/sat.ncu-rep > $O/ncu_saturation.txt 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gsw_sweep_kernel -s 2 -c 1 -f -o $R/sweep python tools/perf_greedy.py > $O/ncu_sweep.log 2>&1
python tools/ncu_summary.py $R/sweep.ncu-rep > $O/ncu_sweep.txt 2>&1
timeout 300 python tools/perf_sat.py 1000000 12 > $O/perf_sat.json 2>/dev/null
timeout 300 python tools/perf_greedy.py > $O/perf_greedy.json 2>/dev/null
timeout 300 python tools/perf_sizer_full.py 1.0 > $O/perf_sizer_full.json 2>/dev/null
timeout 600 python tools/cfg5_ingest.py > $O/cfg5_ingest.json 2> $O/cfg5_ingest.err
tail -3 $O/pytest_gpu.log; cat $O/smoke.log | tail -2; tail -2 $O/bench_1gpu.err; python - <<PY
import json
for f in ('bench_1gpu','bench_reference_arm'):
    try:
        d=json.load(open('$O/'+f+'.json'))
        print(f, d.get('value'), d.get('ms_per_step'), json.dumps(d.get('solver_wall_ms'))[:400], json.dumps(d.get('e2e'))[:200], json.dumps(d.get('cpu_baseline'))[:300], json.dumps(d.get('roofline_hbm'))[:200], json.dumps(d.get('clocks')))
    except Exception as e: print(f, 'ERR', e)
PY
ls -la $O
