#!/usr/bin/env python3
"""Hottest SASS instructions (by stall samples) of the first kernel in an .ncu-rep, with their top stall reasons.
usage: ncu_hot.py report.ncu-rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 24
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True).stdout.decode("utf-8", "replace")
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]; idx = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
out = []
for i, r in enumerate(rows[2:]):
    try: n = int(r[5]); s = int(r[4])
    except (ValueError, IndexError): continue
    out.append((i, r[1].strip(), n, s, r[8], {hdr[j][6:]: int(r[j]) for j in idx if r[j] not in ("0", "")}))
tot = sum(o[3] for o in out) or 1; ti = sum(o[2] for o in out)
print("total samples", tot, "warp instructions", ti)
for o in sorted(out, key=lambda x: -x[3])[:top]:
    reasons = sorted(o[5].items(), key=lambda kv: -kv[1])[:2]
    print(f"{o[0]:5d} n={o[2]/1e6:7.1f}M samp={100*o[3]/tot:5.1f}% thr={o[4]:>4s} {o[1][:72]:72s} {reasons}")
