#!/usr/bin/env python3
"""Kernel-only driver for ncu captures: runs the inflate kernel (BAMSCAN_DEBUG_FLAGS=4 forces the warp-per-member kernel, 8 the lane-group kernel; and optionally one full device-resident pass) on the cached
bench file.  usage: python tools/prof_inflate.py [reads] [full]"""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "datafusion-bio-formats_b200"))
import bench, bamscan
reads = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
path, info = bench.ensure_bam(reads, 2, True)
p = bamscan.BamTableProvider(str(path), None, True, bench.TAGS, False, True, 100, None, debug_flags=int(os.environ.get("BAMSCAN_DEBUG_FLAGS", "0")), skip_crc=bool(int(os.environ.get("BAMSCAN_SKIP_CRC", "0"))))
plan = p.scan(None, [], None, target_partitions=1, partition_mode="block_range")
r = plan.bench_inflate(0, 5)
gb = (r["inflated_bytes"] + r["compressed_bytes"]) / 1e9
print("inflate:", r, "GB/s(in+out)=%.1f" % (gb / (r["ms_per_launch"] / 1e3)), "GB/s(out)=%.1f" % (r["inflated_bytes"] / 1e9 / (r["ms_per_launch"] / 1e3)))
if len(sys.argv) > 2:
    print(plan.run_device_resident(0, 2))
