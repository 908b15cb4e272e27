#!/usr/bin/env python
"""Summarise an .ncu-rep (run where ncu is installed; no GPU needed): key metrics, stall reasons, hottest SASS lines."""
import csv, subprocess, sys, io
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed_op_local_ld.sum",
        "smsp__inst_executed_op_local_st.sum", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k); print(f"{k} = {r[i]} {units[i]}")
    print("-- stall samples (pc sampling):")
    st = []
    for i, h in enumerate(hdr):
        if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
            try: st.append((float(r[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError: pass
    tot = sum(v for v, _ in st) or 1
    for v, h in sorted(st, reverse=True)[:8]: print(f"   {v:8.0f} {v / tot:6.1%} {h}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
for idx, rr in enumerate(rows):
    if "Source" in rr and any("Sampling" in c for c in rr): break
hdr = rows[idx]
si, ws, ie = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
data = []
for n, rr in enumerate(rows[idx + 1:]):
    try: data.append((float(rr[ws]), n, rr[si][:90], rr[ie]))
    except (ValueError, IndexError): pass
tot = sum(d[0] for d in data) or 1
print(f"-- hottest SASS (of {tot:.0f} samples, {len(data)} instructions):")
for v, n, s_, e in sorted(data, reverse=True)[:top]: print(f"   {v:6.0f} {v / tot:5.1%} #{n:<6d} x{e:>9s}  {s_}")
