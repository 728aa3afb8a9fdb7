#!/usr/bin/env python3
"""Summarises one kernel of an .ncu-rep into markdown: key raw metrics + hottest CUDA source lines.
usage: ncu_summary.py report.ncu-rep "title" > profiles/xyz.md"""
import csv, io, subprocess, sys
rep, title = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum"]
print(f"# {title}\n")
print(f"source: `{rep}` (ncu --set full --clock-control none --import-source on; one launch)\n")
print("| metric | value | unit |\n|---|---:|---|")
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print(f"| {h} | {v[:70]} | {u} |")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None; data = []
for r in rows:
    if r and r[0] == "Line No":
        hdr = r; ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples"); continue
    if hdr is None or len(r) <= ie or not r[0].strip().isdigit(): continue
    try: data.append((int(r[ie]), int(r[isamp]), int(r[0]), r[1]))
    except ValueError: pass
tot = sum(d[0] for d in data) or 1; tots = sum(d[1] for d in data) or 1
print(f"\n## hottest source lines ({tot} warp instructions attributed, {tots} stall samples)\n")
print("| % inst | % samples | line | source |\n|---:|---:|---:|---|")
for n, s, ln, text in sorted(data, key=lambda x: -x[0])[:18]:
    print(f"| {100*n/tot:.1f} | {100*s/tots:.1f} | {ln} | `{text.strip()[:110].replace('|', '/')}` |")
