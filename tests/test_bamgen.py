"""CPU: the repo's synthetic BAM writer (tools/bamgen.cpp) -- determinism, validity, sortedness, BAI offsets."""
import hashlib
import struct
import subprocess

import pyarrow as pa
import pyarrow.compute as pc

from conftest import build_bamgen, gen_bam


def test_deterministic_across_thread_counts(tmp_path):
    exe = build_bamgen()
    outs = []
    for th in (1, 3):
        o = tmp_path / f"t{th}.bam"
        subprocess.check_call([str(exe), "--mode", "short", "--reads", "40000", "--seed", "9", "--threads", str(th), "--out", str(o), "--bai"], stdout=subprocess.DEVNULL)
        outs.append((hashlib.sha256(o.read_bytes()).hexdigest(), hashlib.sha256((tmp_path / f"t{th}.bam.bai").read_bytes()).hexdigest()))
    assert outs[0] == outs[1]


def test_short_reads_are_valid_sorted_and_shaped(syn_dir):
    from oracle.bam_oracle import OracleBam
    path = gen_bam(syn_dir, "short", 20000, seed=3)
    o = OracleBam(str(path), tag_fields=["NM", "MD", "AS", "XS", "RG", "MC", "MQ"])
    t = pa.Table.from_batches([o.scan()])
    assert t.num_rows == 20000
    assert len(o.ref_names) == 25 and o.ref_names[0] == "chr1" and o.ref_lens[0] == 248956422
    ref_idx = [o.ref_names.index(c) for c in t["chrom"].to_pylist()]
    keys = list(zip(ref_idx, t["start"].to_pylist()))
    assert keys == sorted(keys), "coordinate sorted"
    assert set(pc.utf8_length(t["sequence"]).to_pylist()) == {150}
    nl = pc.utf8_length(t["name"]).to_pylist()
    assert min(nl) >= 28 and max(nl) <= 36
    assert t["NM"].null_count == 0 and t["RG"].to_pylist()[0] == "rg1"
    assert set(t["flags"].to_pylist()) <= {99, 147, 83, 163, 99 | 1024, 147 | 1024, 83 | 1024, 163 | 1024}
    assert all(0 <= q <= 60 for q in t["mapping_quality"].to_pylist())
    assert t["end"].null_count == 0            # every mapped read has a reference-consuming CIGAR (no unpinned rows)


def test_long_reads_span_blocks(syn_dir):
    from oracle.bam_oracle import OracleBam
    path = gen_bam(syn_dir, "long", 300, seed=5, unmapped=7)
    o = OracleBam(str(path), tag_fields=["NM", "MD", "MM", "ML"])
    b, st = o.scan(want_stats=True)
    assert b.num_rows == 307 and st["blocks"] > 60          # ~20 KB records over 64 KiB blocks: most records straddle a seam
    assert b.column("chrom").null_count == 7 and b.column("ML").type == pa.list_(pa.field("item", pa.uint8(), True))
    assert max(pc.utf8_length(b.column("sequence")).to_pylist()) > 20000


def test_bai_linear_offsets_are_record_starts(syn_dir):
    import bench
    from oracle.bam_oracle import OracleBam
    path = gen_bam(syn_dir, "short", 20000, seed=4, bai=True)
    offs = bench.bai_linear_offsets(type(path)(str(path) + ".bai"))
    o = OracleBam(str(path))
    voff, _idx, nrec = o.index_records(stride=1)
    starts = set(int(v) for v in voff)
    assert nrec == 20000 and len(offs) > 100
    assert all(v in starts for v in offs)


def test_deflate_sim_host_model_runs(syn_dir, tmp_path):
    """tools/deflate_sim.cpp (host model of the write path's LZ77 stage, the tool the compressor's variants were sized with)
    builds and prints its table: the committed parser stays within 1.25 x of zlib level 6 on the synthetic stream and the
    two-way buckets are smaller than one-way ones."""
    import re, struct, subprocess, zlib
    from conftest import ROOT, gen_bam
    exe = tmp_path / "deflate_sim"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", str(exe), str(ROOT / "tools" / "deflate_sim.cpp"), "-lz"])
    data = gen_bam(syn_dir, "short", 20000, seed=5).read_bytes()
    off, out = 0, []
    while off < len(data):
        bs = struct.unpack_from("<H", data, off + 16)[0] + 1
        out.append(zlib.decompress(data[off + 18: off + bs - 8], -15)); off += bs
    raw = tmp_path / "s.raw"
    raw.write_bytes(b"".join(out))
    txt = subprocess.check_output([str(exe), str(raw), "40"]).decode()
    rows = {m.group(1).strip(): float(m.group(2)) for m in re.finditer(r"^(.*?)\s+\d+ bytes\s+([0-9.]+) x zlib-6$", txt, re.M)}
    base = next(v for k, v in rows.items() if k.startswith("committed"))
    assert 0.95 < base < 1.25, txt
    assert rows["4-way buckets (same memory)"] < base < rows["one candidate per slot (the kernel before the last change)"]
