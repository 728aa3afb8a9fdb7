#!/usr/bin/env python3
"""Aggregates an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line.
usage: ncu_lines.py dump.csv [top_n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
data = []
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= ie or not r[0].strip().isdigit():
        continue
    try:
        data.append((int(r[ie]), int(r[isamp]), int(r[0]), r[1]))
    except ValueError:
        pass
tot = sum(d[0] for d in data) or 1
tots = sum(d[1] for d in data) or 1
print(f"total warp-instructions {tot}, samples {tots}")
for n, s, ln, src in sorted(data, key=lambda x: -x[0])[:top]:
    print(f"{100*n/tot:5.1f}% inst {100*s/tots:5.1f}% samp  L{ln:<4d} {src.strip()[:120]}")
