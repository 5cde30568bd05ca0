#!/usr/bin/env python
"""Top stall sites from an `ncu -i X.ncu-rep --page source --csv` export (run in the build container, no GPU).
    python scripts/ncu_src_top.py gpurun_out/x_src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
num = lambda r, h: int(float(r[idx[h]] or 0))
tot = sum(num(r, "# Samples") for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("total samples", tot)
agg = {h: sum(num(r, h) for r in data) for h in stalls}
print("by reason:", {k: f"{100*v/tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:n]:
    s = num(r, "# Samples")
    st = sorted(((h, num(r, h)) for h in stalls), key=lambda kv: -kv[1])[:2]
    print(r[0][-5:], f"{100*s/tot:5.1f}%", f"{num(r, 'Instructions Executed'):>9}", r[1][:64].ljust(64), st)
