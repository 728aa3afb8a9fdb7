"""GPU parity: libbamscan.so (through the C ABI) vs the CPU oracle, bit-exact (RecordBatch equality).

Mirrors the reference's own integration tests where they exist:
  bio-format-bam/tests/indexed_read_test.rs (row counts 421 = 160+159+102), indexed_read_large_test.rs (4277),
  projection_pushdown_test.rs (column order follows the projection, COUNT(*) via empty projection, tail batch kept),
  tag_tests.rs (tag column types / nulls), sam_read_test.rs value pins (MAPQ 255, name "*", TLEN sign).
"""
import collections

import pyarrow as pa
import pytest

from conftest import GOLDEN, gen_bam

pytestmark = pytest.mark.gpu


def _oracle(path, **kw):
    from oracle.bam_oracle import OracleBam
    return OracleBam(str(path), **kw)


# Every test of this module runs once per inflate kernel (debug_flags): bit 4 forces the CTA-per-member kernel
# (kernels_inflate_cta.cuh, what every launch of <= 16384 members uses), bit 2 the warp-per-member kernel (its retry path),
# bit 3 the lane-group kernel (full waves of a big scan).
_FORCE_INFLATE = 0


@pytest.fixture(autouse=True, params=[16, 4, 8], ids=["cta_per_member", "warp_per_member", "lane_group"])
def inflate_kernel(request):
    global _FORCE_INFLATE
    _FORCE_INFLATE = request.param
    yield request.param
    _FORCE_INFLATE = 0


def _provider(path, **kw):
    import bamscan
    zero_based = kw.pop("zero_based", True)
    kw["debug_flags"] = kw.get("debug_flags", 0) | _FORCE_INFLATE
    kw.setdefault("index_path", "")      # these tests pin the sequential path (physical_exec.rs:371-598); indexed scans: test_gpu_indexed.py
    return bamscan.BamTableProvider(str(path), None, zero_based, kw.pop("tag_fields", None), kw.pop("binary_cigar", False),
                                    kw.pop("infer_tag_types", True), kw.pop("infer_tag_sample_size", 100),
                                    kw.pop("tag_type_hints", None), **kw)


def _assert_tables_equal(got: pa.Table, want: pa.RecordBatch, ctx=""):
    want_t = pa.Table.from_batches([want])
    assert got.num_rows == want_t.num_rows, f"{ctx}: rows {got.num_rows} != {want_t.num_rows}"
    assert got.schema.names == want_t.schema.names, ctx
    for name in want_t.schema.names:
        g = got[name].combine_chunks()
        w = want_t[name].combine_chunks()
        assert g.type == w.type, f"{ctx}: column {name} type {g.type} != {w.type}"
        if not g.equals(w):
            for i in range(len(w)):
                if g[i] != w[i]:
                    raise AssertionError(f"{ctx}: column {name} row {i}: got {g[i]!r} want {w[i]!r}")
            raise AssertionError(f"{ctx}: column {name} differs (null masks?)")


FIXTURES = [
    ("multi_chrom.bam", None, 421),
    ("multi_chrom_large.bam", None, 4277),
    ("nanopore_custom_tags.bam", ["pa", "ns", "NM", "ts", "de", "MD", "tp"], 20),
    ("bam_with_tags.bam", ["XT", "NM", "MD", "RG", "UQ"], 14),
    ("10x_pbmc_tags.bam", ["RE", "ts", "CB", "NH", "xf"], 10),
    ("no_coor_only.bam", ["CB", "CR"], 2),
]


@pytest.mark.parametrize("name,tags,rows", FIXTURES)
@pytest.mark.parametrize("zero_based", [True, False])
def test_fixture_full_projection(name, tags, rows, zero_based):
    from oracle.bam_oracle import schema_equal
    path = GOLDEN / name
    o = _oracle(path, zero_based=zero_based, tag_fields=tags)
    p = _provider(path, zero_based=zero_based, tag_fields=tags)
    assert schema_equal(p.schema(), o.schema), f"{p.schema()} vs {o.schema}"
    plan = p.scan(None, [], None)
    assert plan.output_partition_count() == 1
    got = plan.collect()
    assert got.num_rows == rows
    _assert_tables_equal(got, o.scan(), name)


def test_multi_chrom_counts_and_pins():
    # indexed_read_test.rs:76-77,108,121 ; SURVEY App. C first/last record
    p = _provider(GOLDEN / "multi_chrom.bam")
    t = p.scan(None, [], None).collect()
    c = collections.Counter(t["chrom"].to_pylist())
    assert (c["chr1"], c["chr2"], c["chrX"]) == (160, 159, 102)
    first = t.slice(0, 1).to_pylist()[0]
    assert first["name"] == "61CC3AAXX100125:7:10:9729:16571" and first["start"] == 55000103 and first["end"] == 55000179
    assert first["cigar"] == "76M" and first["flags"] == 99 and first["mapping_quality"] == 99 and first["template_length"] == 228
    last = t.slice(t.num_rows - 1, 1).to_pylist()[0]
    assert last["cigar"] == "12S89M" and last["template_length"] == -88 and last["chrom"] == "chrX"


@pytest.mark.parametrize("projection", [[1, 2, 3, 6, 4], [9, 0], [11], [5, 10], [], [12, 2], [13], [0, 0]])
def test_projection_order(projection):
    # projection_pushdown_test.rs:115-126 (order follows SELECT), :157-169 (empty projection keeps the count)
    path = GOLDEN / "multi_chrom_large.bam"
    tags = ["NM", "MD"]
    o = _oracle(path, tag_fields=tags)
    p = _provider(path, tag_fields=tags)
    plan = p.scan(projection, [], None)
    if not projection:
        batches = plan.collect()
        assert sum(b.num_rows for b in batches) == 4277
        assert all(b.num_columns == 0 for b in batches)
        return
    got = plan.collect()
    want = o.scan(projection=projection)
    assert got.schema.names == [o.schema.names[i] for i in projection]
    if len(set(projection)) == len(projection):
        _assert_tables_equal(got, want, str(projection))
    else:
        assert got.num_rows == want.num_rows and got.column(0).equals(got.column(1))


def test_binary_cigar():
    path = GOLDEN / "multi_chrom.bam"
    o = _oracle(path, binary_cigar=True)
    p = _provider(path, binary_cigar=True)
    got = p.scan([5, 0], [], None).collect()
    _assert_tables_equal(got, o.scan(projection=[5, 0]), "binary cigar")
    assert got.schema.field("cigar").type == pa.binary()
    assert got.schema.metadata[b"bio.bam.binary_cigar"] == b"true"


def test_tag_type_hints_and_inference_off():
    path = GOLDEN / "nanopore_custom_tags.bam"
    kw = dict(tag_fields=["pa", "ns", "ts"], infer_tag_types=False, tag_type_hints=["pa:B:i", "ns:i", "ts:i"])
    o = _oracle(path, **kw)
    p = _provider(path, **kw)
    from oracle.bam_oracle import schema_equal
    assert schema_equal(p.schema(), o.schema)
    _assert_tables_equal(p.scan(None, [], None).collect(), o.scan(), "hints")


def test_batch_rows_slicing_keeps_tail():
    # physical_exec.rs:545-593: batches of <= batch_size rows, partial tail batch emitted (projection_pushdown_test.rs:177-188)
    path = GOLDEN / "multi_chrom_large.bam"
    p = _provider(path, batch_rows=1000)
    batches = list(p.scan(None, [], None).execute(0))
    assert [b.num_rows for b in batches] == [1000, 1000, 1000, 1000, 277]
    o = _oracle(path)
    _assert_tables_equal(pa.Table.from_batches(batches), o.scan(), "sliced")


@pytest.mark.parametrize("chunk,seg", [(0, 0), (1 << 20, 4096), (300000, 1024), (70000, 256)])
def test_synthetic_short_chunks_and_carry(syn_dir, chunk, seg):
    path = gen_bam(syn_dir, "short", 20000, seed=3)
    tags = ["NM", "MD", "AS", "RG"]
    o = _oracle(path, tag_fields=tags)
    p = _provider(path, tag_fields=tags, chunk_inflated_bytes=chunk, segment_bytes=seg)
    plan = p.scan(None, [], None)
    got = plan.collect()
    _assert_tables_equal(got, o.scan(), f"chunk={chunk} seg={seg}")


@pytest.mark.parametrize("chunk", [0, 1 << 20, 200000])
def test_synthetic_long_reads(syn_dir, chunk):
    # config 5: records spanning many BGZF blocks, long CIGARs, MM:Z / ML:B:C
    path = gen_bam(syn_dir, "long", 300, seed=5, unmapped=7)
    tags = ["NM", "MD", "MM", "ML"]
    o = _oracle(path, tag_fields=tags)
    p = _provider(path, tag_fields=tags, chunk_inflated_bytes=chunk)
    got = p.scan(None, [], None).collect()
    _assert_tables_equal(got, o.scan(), f"long chunk={chunk}")


@pytest.mark.parametrize("nparts", [2, 3, 8])
@pytest.mark.parametrize("mode", ["short", "long"])
def test_block_range_partitions_concatenate_to_full_scan(syn_dir, nparts, mode):
    path = gen_bam(syn_dir, mode, 20000 if mode == "short" else 300, seed=7)
    o = _oracle(path)
    p = _provider(path, chunk_inflated_bytes=1 << 20)
    plan = p.scan(None, [], None, target_partitions=nparts, partition_mode="block_range")
    assert plan.output_partition_count() == nparts
    got = plan.collect()
    _assert_tables_equal(got, o.scan(), f"{mode} x{nparts}")


@pytest.mark.parametrize("sequential_only", [0, 64])
def test_boundary_repair_path(syn_dir, sequential_only):
    """bit 6 leaves every disagreeing seam to the sequential repair; without it the parallel rounds (seg_fix_kernel) go first."""
    path = gen_bam(syn_dir, "short", 20000, seed=3)
    o = _oracle(path)
    p = _provider(path, segment_bytes=2048, debug_flags=1 | sequential_only)   # every third segment gets a wrong start
    plan = p.scan([0, 2, 5], [], None)
    got = plan.collect()
    assert plan.last_stats["boundary_repairs"] > 0
    _assert_tables_equal(got, o.scan(projection=[0, 2, 5]), "repair")


def test_crc_corruption_is_detected(tmp_path):
    import bamscan
    src = (GOLDEN / "multi_chrom.bam").read_bytes()
    bad = bytearray(src)
    bad[30000] ^= 0x55     # inside the deflate payload of the first data block
    f = tmp_path / "corrupt.bam"
    f.write_bytes(bytes(bad))
    p = _provider(f)
    with pytest.raises(bamscan.BamScanError):
        p.scan(None, [], None).collect()


def test_not_a_bam(tmp_path):
    import bamscan
    f = tmp_path / "x.bam"
    f.write_bytes(b"\n\n<!DOCTYPE html>" + b"x" * 100)   # what the reference's rev_reads.bam fixture really is
    with pytest.raises(bamscan.BamScanError):
        _provider(f).scan(None, [], None)


@pytest.mark.parametrize("seed", range(6))
def test_corrupt_payload_bytes_never_pass_silently(tmp_path, seed):
    """Fuzz: single corrupted bytes anywhere in the deflate payloads / trailers must surface as an error (CRC32, ISIZE, Huffman,
    distance, overrun ...) -- never as a hang, a crash or silently different rows."""
    import random
    import bamscan
    rng = random.Random(seed)
    src = (GOLDEN / "multi_chrom_large.bam").read_bytes()
    o_rows = 4277
    for k in range(4):
        bad = bytearray(src)
        pos = rng.randrange(3000, len(src) - 28)          # past the header block's first bytes, before the EOF marker
        bad[pos] ^= 1 << rng.randrange(8)
        f = tmp_path / f"fz_{seed}_{k}.bam"
        f.write_bytes(bytes(bad))
        try:
            p = _provider(f)
            t = p.scan([0, 2, 9], [], None).collect()
        except bamscan.BamScanError:
            continue
        # only a flip inside a BGZF header field the scan does not use (MTIME, XFL, OS) may leave the result intact
        assert t.num_rows == o_rows


@pytest.mark.parametrize("name,tags", [("multi_chrom_large.bam", ["NM", "MD", "RG", "OQ", "XT"]), ("nanopore_custom_tags.bam", ["pa", "ns", "NM", "ts", "de", "MD", "tp"]),
                                       ("10x_pbmc_tags.bam", ["RE", "ts", "CB", "NH", "xf"]), ("no_coor_only.bam", ["CB", "CR"])])
def test_long_record_decode_kernel_matches_on_short_records(name, tags):
    """decode_fixed_warp_kernel (a warp per 32 rows, chosen automatically for long-read files) forced on ordinary fixtures."""
    path = GOLDEN / name
    o = _oracle(path, tag_fields=tags)
    p = _provider(path, tag_fields=tags, debug_flags=2)
    _assert_tables_equal(p.scan(None, [], None).collect(), o.scan(), f"warp decode {name}")


@pytest.mark.parametrize("level", [0, 1, 9])
def test_deflate_block_types(syn_dir, level):
    """Stored blocks (level 0), fixed-Huffman-heavy output (level 1) and the densest dynamic codes (level 9)."""
    path = gen_bam(syn_dir, "short", 6000, seed=11, level=level)
    tags = ["NM", "MD", "AS", "RG"]
    o = _oracle(path, tag_fields=tags)
    p = _provider(path, tag_fields=tags, chunk_inflated_bytes=1 << 20)
    _assert_tables_equal(p.scan(None, [], None).collect(), o.scan(), f"level {level}")
