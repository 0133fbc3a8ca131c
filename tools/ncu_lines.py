"""Aggregate a `ncu --page source --csv --print-source cuda,sass` export by CUDA source line: warp instructions executed
(divided by `units`, e.g. the number of models) and the share of the stall samples per line.
Usage: ncu_lines.py export.csv [units=1] [min=1.5]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
mn = float(sys.argv[3]) if len(sys.argv) > 3 else 1.5
agg = collections.OrderedDict(); fname = None; hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split('/')[-1]; continue
    if r[0] == "Line No":
        hdr = {}
        for i, h in enumerate(r): hdr.setdefault(h, i)
        continue
    if hdr is None or len(r) < 10 or not r[0].isdigit(): continue
    try: ie = int(r[hdr["Instructions Executed"]]); sm = int(r[hdr["# Samples"]])
    except Exception: continue
    if ie > 0: agg[(fname, int(r[0]))] = (ie, sm, r[1].strip()[:110])
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print(f"total {tot/units:.1f} per unit, {ts} samples")
for (f, l), (ie, sm, src) in agg.items():
    if ie / units >= mn: print(f"{f[:20]:20s}{l:5d} {ie/units:7.1f} {100*sm/max(ts,1):5.1f}%  {src}")
