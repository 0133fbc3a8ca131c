#!/bin/bash
# Final-build evidence after the saturation-kernel rewrite: full GPU suite, smoke, the bench line (both arms), the ncu launch
# list of the bench command, an `ncu --set full` capture of the saturation kernel (summary + per-source-line instruction
# counts; the reports are too big to travel), the kernel not under ncu, and the configs[4] ingest cycle.  The other dominant
# kernels (sizer_pool, grid, sweep) are unchanged since their final-build captures (profiles/README.md).
TAG=${1:-final2}
O=gpurun_out/$TAG
mkdir -p $O
echo "build: $(cat .git_rev 2>/dev/null)" > $O/build.txt
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 900 python bench.py > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "bench rc=$?" >> $O/bench_1gpu.err
timeout 900 python bench.py --impl reference > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/launches_bench.out 2>&1
R=/tmp/ncu_reps; mkdir -p $R
timeout 600 ncu --set full --import-source on --clock-control none -k regex:saturation_kernel -s 3 -c 1 -f -o $R/sat python tools/perf_sat.py 1000000 4 > $O/ncu_sat.log 2>&1
python tools/ncu_summary.py $R/sat.ncu-rep > $O/ncu_saturation.txt 2>&1
ncu -i $R/sat.ncu-rep --page source --csv --print-source cuda,sass > /tmp/sat_src.csv 2>/dev/null
python tools/ncu_lines.py /tmp/sat_src.csv 1e6 1.0 > $O/ncu_saturation_lines.txt 2>&1
timeout 300 python tools/perf_sat.py 1000000 12 > $O/perf_sat.json 2>/dev/null
timeout 600 python tools/cfg5_ingest.py > $O/cfg5_ingest.json 2> $O/cfg5_ingest.err
tail -3 $O/pytest_gpu.log; tail -2 $O/smoke.log; tail -2 $O/bench_1gpu.err; python - <<PY
import json
for f in ('bench_1gpu','bench_reference_arm'):
    try:
        d=json.load(open('$O/'+f+'.json'))
        print(f, d.get('value'), d.get('ms_per_step'), json.dumps(d.get('solver_wall_ms'))[:300], json.dumps(d.get('e2e'))[:200], json.dumps(d.get('cpu_baseline'))[:200], json.dumps(d.get('roofline_hbm'))[:300], json.dumps(d.get('clocks')))
    except Exception as e: print(f, 'ERR', e)
PY
cat $O/perf_sat.json | head -c 600; echo; head -c 600 $O/cfg5_ingest.json; echo; ls -la $O
