import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "datafusion-bio-formats_b200"))
import subprocess, bamscan
lp = Path("/dev/shm/bamscan_bench/long_200000_s5.bam"); lp.parent.mkdir(parents=True, exist_ok=True)
if not lp.exists():
    subprocess.check_call([str(ROOT / "tools/_build/bamgen"), "--mode", "long", "--reads", "200000", "--seed", "5", "--out", str(lp)], stdout=subprocess.DEVNULL)
p = bamscan.BamTableProvider(str(lp), None, True, ["NM", "MD", "MM", "ML"], False, True, 100, None, index_path="")
plan = p.scan(None, [], None, target_partitions=1, partition_mode="block_range")
plan.run_device_resident(0, 1)   # warm-up: device buffers are allocated here and kept by the handle
print(plan.run_device_resident(0, 2))
