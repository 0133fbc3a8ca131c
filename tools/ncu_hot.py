"""Print the hottest SASS region of an .ncu-rep source page (instruction counts + stall samples)."""
import csv, subprocess, sys
rep = sys.argv[1]; frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
k = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[k]; ix = {h: i for i, h in enumerate(hdr)}
ie, isrc, isamp, ithr = ix["Instructions Executed"], ix["Source"], ix["# Samples"], ix["Avg. Threads Executed"]
body = [r for r in rows[k + 1:] if len(r) > ie and r[ie].isdigit()]
tot = sum(int(r[ie]) for r in body); tots = sum(int(r[isamp]) for r in body)
print("total warp-instr", tot, "samples", tots)
for i, r in enumerate(body):
    if int(r[ie]) > tot * frac:
        print(f"{i:5d} {int(r[ie]):>12d} {100*int(r[isamp])/max(tots,1):5.2f}% thr={r[ithr]:>5s} {r[isrc].strip()[:100]}")
