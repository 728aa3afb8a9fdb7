#!/usr/bin/env python3
"""Config 4 (BAI region query) stage breakdown per partition, for the default kernel choice and with each inflate kernel forced.
usage: python tools/region_trace.py [reads]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "datafusion-bio-formats_b200"))
import bench, bamscan
reads = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
path, info = bench.ensure_bam(reads, 2, True)
flt = [("chrom", "=", ["chr1"]), ("start", "between", [50_000_000, 150_000_000])]
for flags in (0, 4, 8):
    prov = bamscan.BamTableProvider(str(path), None, True, bench.TAGS, False, True, 100, None, debug_flags=flags)
    for tp in (1, 2):
        plan = prov.scan(None, flt, None, target_partitions=tp)
        for p in range(plan.output_partition_count()):
            plan.run_device_resident(p, 1)
            st = plan.run_device_resident(p, 2)
            print(f"flags={flags} tp={tp} part={p} ranges={len(plan.partition_ranges(p))} blocks={st['blocks']} chunks={st['chunks']} rows={st['rows']} "
                  f"total={st['ms_total']:.1f} inflate={st['ms_inflate']:.1f} boundary={st['ms_boundary']:.1f} decode={st['ms_decode']:.1f} launches={st["kernel_launches"]} repairs={st["boundary_repairs"]} seam_mismatches={st["boundary_seam_mismatches"]}", flush=True)
    prov.close()
