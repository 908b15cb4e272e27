#!/usr/bin/env python
"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name: python tools/launch_shares.py launches.csv [skip]"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if len(r) > 10 and r[0].isdigit()]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += float(r[-1].replace(",", ""))
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot / 1e6:.3f} ms total (serialised, cold-cache: compare shares)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {v[1] / tot:6.1%} {v[1] / 1e6:9.3f} ms {v[0]:6d} x {v[1] / v[0] / 1e3:9.1f} us  {k}")
