#!/bin/bash
# kernel shares of one prefill pass (2048 tokens, 36 layers) through the C-ABI mirror: ncu launch list of bench.py's prefill leg
mkdir -p gpurun_out
timeout 800 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv --log-file gpurun_out/launches_prefill.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-plugin-e2e > gpurun_out/ncu_prefill.log 2>&1
tail -1 gpurun_out/ncu_prefill.log | cut -c1-200
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches_prefill.csv", errors="ignore")) if len(r) > 10 and r[0].isdigit()]
# the prefill leg runs after the decode legs: take the LAST pass = launches after the last k_cpy_contig<float,__half> mask cast that precedes >100 non-stream kernels
idx = [i for i, r in enumerate(rows) if "k_cpy_contig" in r[4]]
start = idx[-1] if idx else 0
sel = rows[start:]
agg = collections.OrderedDict()
for r in sel:
    a = agg.setdefault(r[4].split("(")[0], [0, 0.0]); a[0] += 1; a[1] += float(r[-1].replace(",", ""))
tot = sum(v[1] for v in agg.values())
print(f"last prefill pass: {len(sel)} launches, {tot/1e6:.3f} ms (serialised, cold-cache: compare shares)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {v[1]/tot:6.1%} {v[1]/1e6:8.3f} ms {v[0]:5d} x {v[1]/v[0]/1e3:8.1f} us  {k}")
PY
