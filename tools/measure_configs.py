#!/usr/bin/env python3
"""Measures the BASELINE.json configs that are not bench.py lines (3: fixed-width projection, 4: BAI region query over
1/2/4/8 partitions, 5: long reads) on ONE GPU: device-resident CUDA-event time and end-to-end wall time.
usage: python tools/measure_configs.py [reads_short] [reads_long]   -> markdown table on stdout"""
import subprocess, sys, time, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "datafusion-bio-formats_b200"))
import bench, bamscan

reads_short = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
reads_long = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
path, info = bench.ensure_bam(reads_short, 2, True)
rows = []

def run(label, provider, projection, filters, tp, mode="reference"):
    plan = provider.scan(projection, filters, None, target_partitions=tp, partition_mode=mode)
    n = plan.output_partition_count()
    for rep in range(2):   # second repetition is reported (buffers warm)
        dev_ms, nrows, infl, arrow = 0.0, 0, 0, 0
        for p in range(n):
            st = plan.run_device_resident(p, 1)
            dev_ms = max(dev_ms, st["ms_total"]) if tp > 1 else dev_ms + st["ms_total"]
            nrows += st["rows"]; infl += st["inflated_bytes"]; arrow += st["arrow_bytes"]
        t0 = time.perf_counter(); e2e_rows = 0
        for p in range(n):
            for b in plan.execute(p):
                e2e_rows += b.num_rows
                del b
        e2e_ms = (time.perf_counter() - t0) * 1e3
    rows.append((label, n, nrows, infl / 1e9, arrow / 1e9, dev_ms, e2e_ms))
    print(f"[measure] {label}: partitions={n} rows={nrows} dev={dev_ms:.1f} ms e2e={e2e_ms:.1f} ms", file=sys.stderr, flush=True)

prov = bamscan.BamTableProvider(str(path), None, True, bench.TAGS, False, True, 100, None, index_path="")
run("2 full projection, sequential", prov, None, [], 1, "block_range")
run("3 chrom,start,end,mapq,flags", prov, [1, 2, 3, 6, 4], [], 1, "block_range")
run("count(*) (empty projection)", prov, [], [], 1, "block_range")
prov.close()
provi = bamscan.BamTableProvider(str(path), None, True, bench.TAGS, False, True, 100, None)
flt = [("chrom", "=", ["chr1"]), ("start", "between", [50_000_000, 150_000_000])]
for tp in (1, 2, 4, 8):
    run(f"4 chr1:50M-150M, {tp} partition(s); dev = slowest partition", provi, None, flt, tp)
run("4 chr1:100.0M-101.0M (narrow)", provi, None, [("chrom", "=", ["chr1"]), ("start", "between", [100_000_000, 101_000_000])], 1)
provi.close()
exe = ROOT / "tools" / "_build" / "bamgen"
lp = path.parent / f"long_{reads_long}_s5.bam"
if not lp.exists():
    subprocess.check_call([str(exe), "--mode", "long", "--reads", str(reads_long), "--seed", "5", "--out", str(lp)], stdout=subprocess.DEVNULL)
provl = bamscan.BamTableProvider(str(lp), None, True, ["NM", "MD", "MM", "ML"], False, True, 100, None, index_path="")
run("5 long reads (10 kb), full projection", provl, None, [], 1, "block_range")
print("| config | partitions | rows | inflated GB | Arrow GB | device ms | reads/s (device) | inflated GB/s (device) | e2e ms | reads/s (e2e) |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for label, n, nrows, infl, arrow, dev, e2e in rows:
    print(f"| {label} | {n} | {nrows} | {infl:.2f} | {arrow:.2f} | {dev:.1f} | {nrows / dev * 1e3 / 1e6:.1f} M | {infl / dev * 1e3:.1f} | {e2e:.1f} | {nrows / e2e * 1e3 / 1e6:.1f} M |")
