"""Small end-to-end run for compute-sanitizer: fixtures (sequential + indexed), a tiny synthetic multi-chunk file, long reads."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "datafusion-bio-formats_b200")); sys.path.insert(0, str(ROOT / "tests"))
import bamscan
from conftest import gen_bam
G = ROOT / "tests" / "golden"
rows = 0
for name, tags in [("multi_chrom.bam", ["NM", "MD", "RG"]), ("nanopore_custom_tags.bam", ["pa", "ns", "de"]), ("no_coor_only.bam", ["CB"])]:
    for idx in ("", None):
        p = bamscan.BamTableProvider(str(G / name), None, True, tags, False, True, 100, None, index_path=idx)
        plan = p.scan(None, [("mapping_quality", ">=", [0])], None, target_partitions=3)
        t = plan.collect()
        rows += t.num_rows if hasattr(t, "num_rows") else 0
tmp = Path("/tmp/sanitize"); tmp.mkdir(exist_ok=True)
f = gen_bam(tmp, "short", 6000, seed=3)
p = bamscan.BamTableProvider(str(f), None, True, ["NM", "MD"], False, True, 100, None, chunk_inflated_bytes=300000, segment_bytes=1024)
rows += p.scan(None, [], None, target_partitions=2, partition_mode="block_range").collect().num_rows
f = gen_bam(tmp, "long", 60, seed=5)
p = bamscan.BamTableProvider(str(f), None, True, ["NM", "ML", "MM"], False, True, 100, None, chunk_inflated_bytes=400000)
rows += p.scan(None, [], None).collect().num_rows
print("rows", rows)
