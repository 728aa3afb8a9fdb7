import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "datafusion-bio-formats_b200"))
import bench, bamscan
reads = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
path, info = bench.ensure_bam(reads, 2, True)
p = bamscan.BamTableProvider(str(path), None, True, bench.TAGS, False, True, 100, None)
plan = p.scan(None, [], None, target_partitions=1, partition_mode="block_range")
plan.run_device_resident(0, 1)   # warm-up: device buffers are allocated here and kept by the handle
print(plan.run_device_resident(0, 2))
