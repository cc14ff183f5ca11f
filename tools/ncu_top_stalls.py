#!/usr/bin/env python
"""Top stalled SASS instructions of an ncu source page:  ncu -i X.ncu-rep --page source --csv | python tools/ncu_top_stalls.py [N]"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
tot = sum(int(r[col["# Samples"]] or 0) for r in data)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("total samples", tot, "instructions", len(data))
agg = {s: sum(int(r[col[s]] or 0) for r in data) for s in stalls}
print("by reason:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for idx, r in sorted(enumerate(data), key=lambda ir: -int(ir[1][col["# Samples"]] or 0))[:n]:
    s = int(r[col["# Samples"]] or 0)
    top = sorted(((int(r[col[k]] or 0), k) for k in stalls), reverse=True)[:2]
    print("%5d %5.1f%%  #%4d %-70s %s" % (s, 100.0 * s / tot, idx, r[col["Source"]][:70], " ".join("%s=%d" % (k[6:], v) for v, k in top if v)))
