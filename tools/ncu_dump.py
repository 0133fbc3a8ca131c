"""Dump a slice of the ncu source page: instr count, samples, wait/math/selected stalls, SASS.  usage: rep start end"""
import csv, subprocess, sys
rep, a, b = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
k = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[k]; ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[k + 1:] if len(r) > ix["Instructions Executed"] and r[ix["Instructions Executed"]].isdigit()]
for i in range(a, min(b, len(body))):
    r = body[i]
    print(i, r[ix["Instructions Executed"]].rjust(9), r[ix["# Samples"]].rjust(5), "wait", r[ix["stall_wait"]].rjust(4), "math", r[ix["stall_math"]].rjust(4),
          "sel", r[ix["stall_selected"]].rjust(4), "nsel", r[ix["stall_not_selected"]].rjust(4), r[ix["Source"]].strip()[:64])
