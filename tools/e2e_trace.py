import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "datafusion-bio-formats_b200"))
import bench, bamscan
reads = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
path, info = bench.ensure_bam(reads, 2, True)
p = bamscan.BamTableProvider(str(path), None, True, bench.TAGS, False, True, 100, None)
plan = p.scan(None, [], None, target_partitions=1, partition_mode="block_range")
for rep in range(3):
    t0 = time.perf_counter(); n = 0
    for b in plan.execute(0):
        n += b.num_rows
        del b
    dt = time.perf_counter() - t0
    print(f"rep {rep}: {n} rows in {dt*1e3:.1f} ms -> {n/dt/1e6:.1f} M reads/s", plan.last_stats["d2h_bytes"]/dt/1e9, "GB/s d2h", file=sys.stderr)
