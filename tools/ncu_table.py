#!/usr/bin/env python
"""One line per kernel of an .ncu-rep (first launch of each kernel name): duration, DRAM bytes and GB/s, % of DRAM / tensor / issue peak, registers.
    python tools/ncu_table.py rep > profiles/...   (runs where ncu is installed; no GPU needed)"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr = rows[0]
col = lambda n: hdr.index(n) if n in hdr else None
C = {k: col(k) for k in ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                         "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
                         "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active"]}
units = rows[1]
def num(r, k):
    i = C[k]
    if i is None: return 0.0
    try: v = float(r[i].replace(",", ""))
    except ValueError: return 0.0
    u = units[i].lower()
    if k.startswith("dram__bytes"): v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    if k.startswith("gpu__time"): v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(u, 1)
    return v
seen = set()
print(f"{'kernel':46s} {'us':>9s} {'dram MB':>9s} {'GB/s':>8s} {'dram%':>6s} {'tensor%':>8s} {'issue%':>7s} {'regs':>5s} {'grid x block':>14s}")
for r in rows[2:]:
    name = r[C["Kernel Name"]].split("(")[0].replace("void ", "").replace("b200::", "")
    if name in seen: continue
    seen.add(name)
    us = num(r, "gpu__time_duration.sum"); mb = (num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum")) / 1e6
    print(f"{name[:46]:46s} {us:9.1f} {mb:9.1f} {mb / us * 1e3 if us else 0:8.0f} {num(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
          f"{num(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):8.1f} {num(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):7.1f} "
          f"{int(num(r, 'launch__registers_per_thread')):5d} {int(num(r, 'launch__grid_size')):7d} x {int(num(r, 'launch__block_size')):4d}")
