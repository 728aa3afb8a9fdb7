#!/usr/bin/env python3
"""Summarise an ncu report of one kernel: headline counters + executed instructions / stall samples per code region.
usage: python tools/ncu_regions.py report.ncu-rep [bucket_bytes=0x200] [min_pct=1.0]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; bucket = int(sys.argv[2], 0) if len(sys.argv) > 2 else 0x200; minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[-1]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
for i, h in enumerate(hdr):
    if h in want or ("issue_stalled" in h and "per_issue_active" in h and float(vals[i] or 0) > 0.05): print(f"{h:90s} {rows[1][i]:12s} {vals[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
ia, isrc, iex, ismp, ith = (hdr.index(x) for x in ("Address", "Source", "Instructions Executed", "# Samples", "Avg. Threads Executed"))
base = int(data[0][ia], 16)
tot = sum(int(r[iex]) for r in data); tots = sum(int(r[ismp]) for r in data)
print(f"total warp instructions {tot}  samples {tots}")
agg = collections.OrderedDict()
for r in data:
    k = (int(r[ia], 16) - base) // bucket
    a = agg.setdefault(k, [0, 0, 0.0, 0]); a[0] += int(r[iex]); a[1] += int(r[ismp]); a[2] += float(r[ith] or 0) * int(r[iex]); 
for k, (e, s, th, _) in agg.items():
    if e > tot * minpct / 100 or s > tots * minpct / 100: print(f"{k*bucket:#7x}  inst {100*e/tot:5.1f}%  samples {100*s/tots:5.1f}%  avg threads {th/max(e,1):4.1f}")
if len(sys.argv) > 4:
    lo, hi = int(sys.argv[4], 0), int(sys.argv[5], 0)
    for r in data:
        off = int(r[ia], 16) - base
        if lo <= off < hi: print(f"{off:#7x} {r[iex]:>10s} {r[ith]:>3s} {r[ismp]:>7s}  {r[isrc].strip()[:80]}")
