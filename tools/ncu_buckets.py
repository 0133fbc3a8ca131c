"""Aggregate an .ncu-rep source page into address buckets: warp instructions, stall samples, FP64 share, average active threads."""
import csv, subprocess, sys
rep = sys.argv[1]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
k = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[k]; ix = {h: i for i, h in enumerate(hdr)}
ie, isrc, isamp, ithr = ix["Instructions Executed"], ix["Source"], ix["# Samples"], ix["Avg. Threads Executed"]
body = [r for r in rows[k + 1:] if len(r) > ie and r[ie].isdigit()]
tot = sum(int(r[ie]) for r in body); tots = sum(int(r[isamp]) for r in body)
print("total warp-instr", tot, "samples", tots)
for b in range(0, len(body), B):
    ch = body[b:b + B]
    n = sum(int(r[ie]) for r in ch); s = sum(int(r[isamp]) for r in ch)
    if n < tot * 0.01 and s < tots * 0.01:
        continue
    f64 = sum(int(r[ie]) for r in ch if r[isrc].strip().split()[0] in ("DFMA", "DMUL", "DADD") or (r[isrc].strip().startswith("@") and r[isrc].split()[1] in ("DFMA", "DMUL", "DADD")))
    thr = sum(float(r[ithr]) * int(r[ie]) for r in ch) / max(n, 1)
    print(f"[{b:5d},{b+B:5d}) instr {100*n/tot:5.1f}%  samples {100*s/max(tots,1):5.1f}%  fp64 share {100*f64/max(n,1):4.0f}%  thr {thr:4.1f}  first: {ch[0][isrc].strip()[:50]}")
