#!/usr/bin/env python
"""Aggregate the warp-stall samples of an .ncu-rep by CUDA source line: python tools/ncu_by_line.py rep [top] [kernel-regex]
(needs a capture made with --import-source on and a build with -lineinfo; runs where ncu is installed, no GPU needed)."""
import csv, io, re, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg, inst, text = defaultdict(float), defaultdict(float), {}
path = None; hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": path = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; ws = r.index("Warp Stall Sampling (All Samples)"); ie = r.index("Instructions Executed"); continue
    if hdr is None or r[0] in ("Function Name",): continue
    try: line = int(r[0])
    except ValueError: continue
    key = (path, line)
    if r[1]: text[key] = r[1].strip()[:110]
    try: agg[key] += float(r[ws] or 0); inst[key] += float(r[ie] or 0)
    except (ValueError, IndexError): pass
tot = sum(agg.values()) or 1
print(f"total samples {tot:.0f}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{v:8.0f} {v / tot:6.1%}  inst {inst[k]:12.0f}  {k[0]}:{k[1]:<5d} {text.get(k, '')}")
