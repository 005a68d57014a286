#!/usr/bin/env python
"""Top stalled SASS instructions of a `ncu --set full --import-source on` capture:
   ncu -i X.ncu-rep --page source --csv > src.csv; python scripts/ncu_top_stalls.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
idx = {k: i for i, k in enumerate(hdr)}
data = []
for r in rows[h + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        data.append((int(r[idx["# Samples"]]), r))
    except ValueError:
        continue
tot = sum(s for s, _ in data)
print(f"{tot} samples, {len(data)} instructions")
for pos, (s, r) in enumerate(data):
    r.append(pos)
for s, r in sorted(data, key=lambda x: -x[0])[:n]:
    print(f"{s:6d} {100 * s / tot:5.1f}%  #{r[-1]:5d}  {r[idx['Source']][:110]}")
