#!/usr/bin/env python
"""Per-kernel table from an .ncu-rep (ncu --set full): duration, DRAM bytes and throughput, tensor-pipe activity.
usage: python scripts/summarize_ncu.py gpurun_out/prof_minibatch.ncu-rep [--md]"""
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
md = "--md" in sys.argv
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]


def col(name):
    return hdr.index(name) if name in hdr else None


c_name, c_t = col("Kernel Name"), col("gpu__time_duration.sum")
c_r, c_w = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
c_dp = col("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
c_tp = col("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") or col(
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active")
c_ia = col("smsp__issue_active.avg.pct_of_peak_sustained_active")
c_grid, c_reg = col("launch__grid_size"), col("launch__registers_per_thread")


def tobytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def tous(v, u):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}.get(u, 1)


tot = 0.0
out = []
for r in rows[2:]:
    name = re.sub(r"^void ", "", re.sub(r"\(.*", "", r[c_name])).replace("mpx::", "")
    t = tous(r[c_t], units[c_t])
    rb, wb = tobytes(r[c_r], units[c_r]), tobytes(r[c_w], units[c_w])
    tot += t
    out.append((name, t, rb, wb, float(r[c_dp]), float(r[c_tp]) if c_tp is not None and r[c_tp] else 0.0,
                float(r[c_ia]) if c_ia is not None else 0.0, r[c_grid], r[c_reg]))
sep = " | " if md else "  "
head = ["#", "kernel", "us", "share%", "MB rd", "MB wr", "GB/s", "dram%", "tensor%", "issue%", "grid", "regs"]
if md:
    print("| " + " | ".join(head) + " |")
    print("|" + "---|" * len(head))
for i, (n, t, rb, wb, dp, tp, ia, g, rg) in enumerate(out):
    cells = [str(i), n[:46], f"{t:.1f}", f"{100 * t / tot:.1f}", f"{rb / 1e6:.1f}", f"{wb / 1e6:.1f}",
             f"{(rb + wb) / t / 1e3:.0f}", f"{dp:.0f}", f"{tp:.0f}", f"{ia:.0f}", g, rg]
    print(("| " + " | ".join(cells) + " |") if md else sep.join(c.ljust(w) for c, w in zip(cells, [3, 46, 8, 6, 8, 8, 6, 5, 7, 6, 8, 4])))
print(f"total {tot:.1f} us over {len(out)} kernels")
