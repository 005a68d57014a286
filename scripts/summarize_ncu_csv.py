#!/usr/bin/env python
"""Per-kernel table from an `ncu --metrics ... --csv --log-file X.csv` capture (long format: one row per kernel x metric).
usage: python scripts/summarize_ncu_csv.py gpurun_out/mb_metrics.csv [--md] [--group]"""
import csv
import re
import sys
from collections import OrderedDict, defaultdict

path = sys.argv[1]
md, group = "--md" in sys.argv, "--group" in sys.argv
lines = [l for l in open(path) if l.startswith('"')]
rows = list(csv.DictReader(lines))
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3,
        "usecond": 1, "msecond": 1e3, "second": 1e6}
k = OrderedDict()
for r in rows:
    d = k.setdefault(r["ID"], {"name": re.sub(r"^void ", "", re.sub(r"\(.*", "", r["Kernel Name"])).replace("mpx::", ""),
                               "grid": r["Grid Size"], "block": r["Block Size"]})
    v = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1)
    d[r["Metric Name"]] = v
ks = list(k.values())
tot = sum(d.get("gpu__time_duration.sum", 0.0) for d in ks)


def fields(d, n=1):
    t = d.get("gpu__time_duration.sum", 0.0)
    rb, wb = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
    return [d["name"][:60], f"{n}", f"{t:.1f}", f"{100 * t / tot:.1f}", f"{rb / 1e6:.2f}", f"{wb / 1e6:.2f}",
            f"{(rb + wb) / max(t, 1e-9) / 1e3:.0f}",
            f"{d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0.0) / n:.1f}",
            f"{d.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0.0) / n:.1f}", d["grid"],
            f"{d.get('launch__registers_per_thread', 0) / n:.0f}"]


head = ["kernel", "n", "us", "share%", "MB rd", "MB wr", "GB/s", "tensor%", "issue%", "grid", "regs"]
out = []
if group:
    g = OrderedDict()
    cnt = defaultdict(int)
    for d in ks:
        key = (d["name"], d["grid"])
        cnt[key] += 1
        if key not in g:
            g[key] = dict(d)
        else:
            for m, v in d.items():
                if isinstance(v, float):
                    g[key][m] = g[key].get(m, 0.0) + v
    for key, d in sorted(g.items(), key=lambda kv: -kv[1].get("gpu__time_duration.sum", 0.0)):
        out.append(fields(d, cnt[key]))
else:
    out = [fields(d) for d in ks]
sep = " | " if md else "  "
w = [max(len(head[i]), max(len(o[i]) for o in out)) for i in range(len(head))]
fmt = lambda r: ("| " if md else "") + sep.join(r[i].ljust(w[i]) for i in range(len(r))) + (" |" if md else "")
print(fmt(head))
if md:
    print("|" + "|".join("-" * (x + 2) for x in w) + "|")
for o in out:
    print(fmt(o))
print(f"\ntotal {tot:.1f} us over {len(ks)} kernels")
