"""Top stall sites of one kernel from an ncu report's source page:
   ncu -i X.ncu-rep --page source --csv --kernel-name K --print-source sass | python scripts/ncu_hot_sass.py [N]"""
import csv
import sys

n = int(sys.argv[1]) if len(sys.argv) > 1 else 25
rows = list(csv.reader(sys.stdin))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]
idx = {h: i for i, h in enumerate(H)}
data = [r for r in rows[hdr + 1:] if len(r) == len(H)]
tot = sum(int(r[idx["# Samples"]] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
stall_cols = [h for h in H if h.startswith("stall_") and "Not Issued" not in h]
order = sorted(range(len(data)), key=lambda i: -int(data[i][idx["# Samples"]] or 0))[:n]
for i in sorted(order):
    r = data[i]
    s = int(r[idx["# Samples"]] or 0)
    top = sorted(((int(r[idx[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print("%5d %5.1f%%  #%-4d %-70s %s" % (s, 100.0 * s / max(tot, 1), i, r[idx["Source"]].strip()[:70], " ".join("%s=%d" % (c, v) for v, c in top if v)))
