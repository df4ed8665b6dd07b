"""Hottest SASS instructions (warp stall samples) of a kernel in an ncu report.  usage: ncu_hot.py <rep> [N]"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(out.splitlines()))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1][:80]
        hdr = rows[i + 1]
        j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            body.append(rows[j]); j += 1
        ks = hdr.index("# Samples"); src = hdr.index("Source"); ie = hdr.index("Instructions Executed")
        tot = sum(float(r[ks] or 0) for r in body) or 1
        print(f"== {name}  total samples {tot:.0f}")
        acc = 0
        for pos, r in sorted(enumerate(body), key=lambda pr: -float(pr[1][ks] or 0))[:N]:
            print(f"  #{pos:4d} {r[src].strip()[:70]:70s} samples {float(r[ks]):7.0f} {100*float(r[ks])/tot:5.1f}%  exec {r[ie]}")
        i = j
    else:
        i += 1
