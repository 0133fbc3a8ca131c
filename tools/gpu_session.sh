#!/bin/bash
# closing session: the pool sizer at the bench's own size under ncu (DRAM traffic per launch for profiles/traffic.json), the
# reference arm after the core-count fix, the bench line once more (saturation leg with the flush synchronised), 2 GPUs skipped
O=gpurun_out/final2; mkdir -p $O /tmp/ncu_reps
timeout 900 ncu --set full --clock-control none -k regex:sizer_pool_kernel -s 1 -c 1 -f -o /tmp/ncu_reps/poolfull python tools/perf_sizer_full.py 1.0 > $O/ncu_poolfull.log 2>&1
python tools/ncu_summary.py /tmp/ncu_reps/poolfull.ncu-rep > $O/ncu_sizer_pool_full.txt 2>&1
timeout 600 ncu --set full --clock-control none -k regex:grid_ -s 2 -c 2 -f -o /tmp/ncu_reps/gridfull python tools/perf_grid.py 1.0 > $O/ncu_gridfull.log 2>&1
python tools/ncu_summary.py /tmp/ncu_reps/gridfull.ncu-rep > $O/ncu_grid_full.txt 2>&1
timeout 900 python bench.py --impl reference > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
timeout 900 python bench.py > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "bench rc=$?" >> $O/bench_1gpu.err
grep -E "gpu__time_duration|dram__bytes|pipe_fp64_cycles" $O/ncu_sizer_pool_full.txt $O/ncu_grid_full.txt
python - <<PY
import json
for f in ('bench_1gpu','bench_reference_arm'):
    try:
        d=json.load(open('$O/'+f+'.json'))
        print(f, d.get('value'), d.get('ms_per_step'), json.dumps(d.get('cpu_baseline'))[:260], json.dumps(d.get('roofline_hbm'))[:160])
    except Exception as e: print(f, 'ERR', e)
PY
