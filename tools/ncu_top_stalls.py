#!/usr/bin/env python
"""Top stalled SASS instructions and the dynamic opcode mix of an ncu source page.

    ncu -i X.ncu-rep --page source --csv | python tools/ncu_top_stalls.py [N] [kernel-substring]
"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(sys.stdin))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
want = sys.argv[2] if len(sys.argv) > 2 else ""
# sections: a "Kernel Name" row, then a header row starting with "Address", then data rows
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1] if len(r) > 1 else "", "hdr": None, "data": []}
        sections.append(cur)
    elif r and r[0] == "Address" and cur is not None:
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
for sec in sections:
    if want not in sec["name"] or not sec["data"]:
        continue
    hdr, data = sec["hdr"], sec["data"]
    col = {h: i for i, h in enumerate(hdr)}
    tot = sum(int(r[col["# Samples"]] or 0) for r in data)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    print("==", sec["name"][:100])
    print("total samples", tot, "static instructions", len(data))
    agg = {s: sum(int(r[col[s]] or 0) for r in data) for s in stalls}
    print("by reason:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
    ex = Counter()
    for r in data:
        toks = [t for t in r[col["Source"]].split() if not t.startswith('@')]
        if toks:
            ex[toks[0].split('.')[0]] += int(r[col["Instructions Executed"]] or 0)
    tex = sum(ex.values())
    print("warp instructions executed", tex, " mix:", ", ".join("%s %.1f%%" % (k, 100.0 * v / tex) for k, v in ex.most_common(16)))
    for idx, r in sorted(enumerate(data), key=lambda ir: -int(ir[1][col["# Samples"]] or 0))[:n]:
        s = int(r[col["# Samples"]] or 0)
        top = sorted(((int(r[col[k]] or 0), k) for k in stalls), reverse=True)[:2]
        print("%5d %5.1f%%  #%4d %-70s %s" % (s, 100.0 * s / max(tot, 1), idx, r[col["Source"]][:70], " ".join("%s=%d" % (k[6:], v) for v, k in top if v)))
