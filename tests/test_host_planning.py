"""CPU: host-side planning of the product (schema, tag typing, filter classification, partitions) against the oracle's
restatement and the reference's pins.  No kernels run here."""
import json
from pathlib import Path

import pyarrow as pa
import pytest

from conftest import GOLDEN, gen_bam


def _provider(path, **kw):
    import bamscan
    return bamscan.BamTableProvider(str(path), None, kw.pop("zero_based", True), kw.pop("tag_fields", None), kw.pop("binary_cigar", False),
                                    kw.pop("infer_tag_types", True), kw.pop("infer_tag_sample_size", 100), kw.pop("tag_type_hints", None), **kw)


CASES = [
    ("multi_chrom.bam", dict()),
    ("multi_chrom.bam", dict(zero_based=False, binary_cigar=True)),
    ("multi_chrom_large.bam", dict(tag_fields=["NM", "MD", "RG", "OQ", "XT", "OP"])),
    ("nanopore_custom_tags.bam", dict(tag_fields=["pa", "ns", "ts", "de", "NM", "ML", "zz"])),
    ("nanopore_custom_tags.bam", dict(tag_fields=["pa", "ns"], infer_tag_types=False)),
    ("nanopore_custom_tags.bam", dict(tag_fields=["pa", "ns", "qq"], infer_tag_types=False, tag_type_hints=["pa:B:i", "ns:I", "qq:f"])),
    ("bam_with_tags.bam", dict(tag_fields=["XT", "X0", "X1", "NM"], infer_tag_sample_size=1)),
    ("10x_pbmc_tags.bam", dict(tag_fields=["CB", "CR", "RE", "xf", "ts", "pa", "GX"])),
    ("no_coor_only.bam", dict(tag_fields=["CB", "CR"])),
]


@pytest.mark.parametrize("name,kw", CASES)
def test_schema_equals_oracle(name, kw):
    from oracle.bam_oracle import OracleBam, schema_equal
    okw = dict(kw)
    o = OracleBam(str(GOLDEN / name), **okw)
    p = _provider(GOLDEN / name, **dict(kw))
    assert schema_equal(p.schema(), o.schema), f"\n{p.schema().to_string(show_schema_metadata=True)}\nvs\n{o.schema.to_string(show_schema_metadata=True)}"


def test_schema_shape_pins():
    # table_provider.rs:57-70: names, types, nullability of the 12 core columns
    s = _provider(GOLDEN / "multi_chrom.bam").schema()
    assert s.names == ["name", "chrom", "start", "end", "flags", "cigar", "mapping_quality", "mate_chrom", "mate_start", "sequence", "quality_scores", "template_length"]
    assert [str(t) for t in s.types] == ["string", "string", "uint32", "uint32", "uint32", "string", "uint32", "string", "uint32", "string", "string", "int32"]
    assert [f.nullable for f in s] == [True, True, True, True, False, False, False, True, True, False, False, False]
    assert s.metadata[b"bio.coordinate_system_zero_based"] == b"true" and b"bio.bam.binary_cigar" not in s.metadata


def test_bad_hint_is_rejected():
    import bamscan
    with pytest.raises(bamscan.BamScanError) as e:
        _provider(GOLDEN / "multi_chrom.bam", tag_fields=["NM"], tag_type_hints=["NM:B"])
    assert "requires a subtype" in str(e.value)


def test_projected_plan_schema_follows_projection():
    # projection_pushdown_test.rs:33-61,115-126
    p = _provider(GOLDEN / "multi_chrom.bam", tag_fields=["NM"])
    plan = p.scan([6, 0, 12], [], None)
    s = plan.schema()
    assert s.names == ["mapping_quality", "name", "NM"] and s.metadata == p.schema().metadata
    assert p.scan([], [], None).schema().names == []
    import bamscan
    with pytest.raises(bamscan.BamScanError):
        p.scan([13], [], None)


def test_filter_classification(tmp_path):
    # table_provider.rs:941-962, genomic_filter.rs:120-148, record_filter.rs:285-355
    import shutil
    idx = _provider(GOLDEN / "multi_chrom.bam", tag_fields=["NM", "MD"])
    f = tmp_path / "noindex.bam"
    shutil.copy(GOLDEN / "multi_chrom.bam", f)
    noidx = _provider(f, tag_fields=["NM", "MD"])
    filters = [("chrom", "=", ["chr1"]), ("chrom", "in", ["chr1", "chr2"]), ("start", "between", [1, 2]), ("end", "<=", [5]),
               ("mapping_quality", ">=", [30]), ("flags", "not_in", [4, 8]), ("name", "=", ["x"]), ("name", "<", ["x"]),
               ("NM", ">", [2]), ("MD", "=", ["10"]), ("sequence", "other", []), ("chrom", "other", []), ("cigar", "!=", ["*"])]
    want_common = ["Inexact"] * 7 + ["Unsupported", "Inexact", "Inexact", "Unsupported", "Unsupported", "Inexact"]
    assert noidx.supports_filters_pushdown(filters) == want_common
    assert idx.supports_filters_pushdown(filters) == want_common      # genomic filters are also record-pushable here
    # a shape only the index makes pushable: start NOT BETWEEN is "genomic" by column name (genomic_filter.rs:130-137) and also record-pushable
    assert idx.supports_filters_pushdown([("start", "not_between", [1, 2])]) == ["Inexact"]


@pytest.mark.parametrize("nparts", [1, 2, 3, 5, 8, 64])
def test_block_range_partitions_tile_the_file(syn_dir, nparts):
    path = gen_bam(syn_dir, "short", 20000, seed=11)
    p = _provider(path)
    plan = p.scan(None, [], None, target_partitions=nparts, partition_mode="block_range")
    assert plan.output_partition_count() == nparts
    ranges = [r for i in range(nparts) for r in plan.partition_ranges(i)]
    assert ranges[0]["exact_start"] == 1 and all(r["exact_start"] == 0 for r in ranges[1:])
    for a, b in zip(ranges, ranges[1:]):
        assert a["block_end"] == b["block_begin"] and a["stop_uoff"] == b["first_uoff"]
    assert ranges[-1]["stop_uoff"] == 2 ** 64 - 1
    sizes = [r["coff_end"] - r["coff_begin"] for r in ranges]
    if nparts <= 8:
        assert max(sizes) - min(sizes) <= 2 * 65536 + 28        # balanced by compressed bytes to within a block or two


def test_reference_partition_rule_without_index(syn_dir):
    # table_provider.rs:1097-1114: no index => UnknownPartitioning(1) whatever target_partitions says
    path = gen_bam(syn_dir, "short", 20000, seed=11)
    p = _provider(path)
    assert p.scan(None, [], None, target_partitions=8).output_partition_count() == 1


def test_not_bgzf_and_missing_file(tmp_path):
    import bamscan
    f = tmp_path / "html.bam"
    f.write_bytes(b"\n\n<!DOCTYPE html><html></html>" * 20)
    prov = _provider(f)                       # provider construction tolerates an unreadable header (table_provider.rs:423-426)
    assert len(prov.schema()) == 12
    with pytest.raises(bamscan.BamScanError):
        prov.scan(None, [], None)
    with pytest.raises(bamscan.BamScanError):
        _provider(tmp_path / "nope.bam")


def test_partition_seam_check_rejects_a_wrong_speculative_start():
    """block-range partitions > 0 speculate their first record; the seam check must fail loudly when a partition does
    not start where its predecessor's chain landed (bamscan_check_partition_seams)."""
    import bamscan
    none = 0xFFFFFFFFFFFFFFFF
    ok = [dict(rows=10, first_record_uoff=100, end_chain_uoff=5000), dict(rows=7, first_record_uoff=5000, end_chain_uoff=9000),
          dict(rows=0, first_record_uoff=none, end_chain_uoff=9000), dict(rows=3, first_record_uoff=9000, end_chain_uoff=12000)]
    bamscan.check_partition_seams(ok)
    bad = [dict(ok[0]), dict(ok[1], first_record_uoff=5012)]
    with pytest.raises(bamscan.BamScanError, match="partition 1 starts at inflated offset 5012"):
        bamscan.check_partition_seams(bad)


def test_corrupt_bai_is_an_error_not_a_crash(tmp_path):
    """n_ref from a corrupt .bai must not turn into a multi-GB allocation / an exception across the C ABI."""
    import shutil
    import bamscan
    bam = tmp_path / "x.bam"
    shutil.copy(GOLDEN / "multi_chrom.bam", bam)
    (tmp_path / "x.bam.bai").write_bytes(b"BAI\1" + (0x7fffffff).to_bytes(4, "little") + b"\0" * 16)
    p = bamscan.BamTableProvider(str(bam), None, True, None, False, True, 100, None)
    try:
        plan = p.scan(None, [("chrom", "=", ["chr1"])], None, target_partitions=4)
        assert plan.output_partition_count() == 1        # an unreadable index = no index: the sequential single-partition plan
    except bamscan.BamScanError as e:
        assert "BAI" in str(e)                           # ... or a clean error; never a crash
    p.close()


def test_open_ended_sub_regions_are_bounded_by_the_start_filter(syn_dir):
    """balance_partitions leaves the last sub-region of a split contig open-ended (partition_balancer.rs: end = None); with a pushed
    `start` upper bound the planner must not read that partition to the end of the contig: no partition's block ranges may reach
    beyond what the one-partition plan (whose region carries the bound) reads.  Without the filter bound they do."""
    import bamscan
    path = gen_bam(syn_dir, "short", 200000, seed=9, bai=True)
    p = bamscan.BamTableProvider(str(path))
    lo, hi = 20_000_000, 60_000_000
    flt = [("chrom", "=", ["chr1"]), ("start", "between", [lo, hi])]
    one = p.scan(None, flt, None, target_partitions=1)
    limit = max(r["coff_end"] for r in one.partition_ranges(0))
    for tp in (2, 4, 8):
        plan = p.scan(None, flt, None, target_partitions=tp)
        n = plan.output_partition_count()
        assert n > 1
        ends = [max((r["coff_end"] for r in plan.partition_ranges(i)), default=0) for i in range(n)]
        assert max(ends) <= limit, (tp, ends, limit)
        regions = [g for i in range(n) for g in plan.partition_regions(i)]
        assert any(g["end"] is None and not g["unmapped_tail"] for g in regions)   # the open-ended last piece is still what is ASSIGNED
    # only a lower bound: the last partition legitimately runs to the end of the contig
    open_plan = p.scan(None, [("chrom", "=", ["chr1"]), ("start", ">=", [lo])], None, target_partitions=4)
    ends = [max((r["coff_end"] for r in open_plan.partition_ranges(i)), default=0) for i in range(open_plan.output_partition_count())]
    assert max(ends) > limit


def test_corrupt_bam_files_open_or_fail_cleanly(tmp_path):
    """bamscan_open walks the BSIZE chain and inflates the header on the host: truncations and byte flips anywhere in a file
    (header block, block headers, payload) either open + plan or raise BamScanError -- never a crash of the host process."""
    import random
    import bamscan
    rng, opened, refused = random.Random(3), 0, 0
    srcs = [(GOLDEN / n).read_bytes() for n in ("multi_chrom.bam", "bam_with_tags.bam", "10x_pbmc_tags.bam")]
    for it in range(300):
        b = bytearray(srcs[it % 3])
        mode = rng.randrange(4)
        if mode == 0:
            b = b[:rng.randrange(len(b))]
        elif mode == 1:
            for _ in range(rng.randrange(1, 8)):
                b[rng.randrange(len(b))] = rng.randrange(256)
        elif mode == 2:
            for _ in range(rng.randrange(1, 8)):
                b[rng.randrange(600)] = rng.randrange(256)
        else:
            pos = rng.randrange(0, 40)
            b[pos:pos + 2] = rng.randrange(65536).to_bytes(2, "little")
        p = tmp_path / "f.bam"
        p.write_bytes(bytes(b))
        try:
            pr = bamscan.BamTableProvider(str(p), None, True, None, False, True, 100, None, index_path="")
            plan = pr.scan(None, [], None, target_partitions=4, partition_mode="block_range")
            for i in range(plan.output_partition_count()):
                plan.partition_ranges(i)
            opened += 1
        except bamscan.BamScanError:
            refused += 1
    assert opened > 0 and refused > 0


def _gzi_bounds(gzi_path, thread_num):
    """get_bgzf_partition_bounds (bio-format-fastq/src/physical_exec.rs:140-175), restated: (start_uncomp, end_comp) per partition."""
    import struct
    d = Path(gzi_path).read_bytes()
    n, = struct.unpack_from("<Q", d, 0)
    blocks = [(0, 0)] + [struct.unpack_from("<QQ", d, 8 + 16 * i) for i in range(n)]
    nb, np_ = len(blocks), min(thread_num, len(blocks))
    out, cur = [], 0
    for i in range(np_):
        if cur >= nb:
            break
        nxt = cur + nb // np_ + (1 if i < nb % np_ else 0)
        out.append((blocks[cur][1], None if nxt >= nb else blocks[nxt][0]))
        cur = nxt
    return out


@pytest.mark.parametrize("tp", [1, 2, 3, 4, 7, 10, 16])
def test_fastq_reference_partitions_follow_the_gzi(tp):
    """FASTQ + companion .gzi: the reference cuts the file into runs of whole BGZF blocks by block count
    (detect_local_strategy, physical_exec.rs:94-116); the plan must expose exactly those bounds."""
    import bamscan
    path = GOLDEN / "fastq" / "sample.fastq.bgz"
    p = bamscan.FastqTableProvider(str(path))
    plan = p.scan(None, None, None, target_partitions=tp, partition_mode="reference")
    want = _gzi_bounds(str(path) + ".gzi", tp)
    assert plan.output_partition_count() == len(want)
    for i, (start_u, end_c) in enumerate(want):
        (r,) = plan.partition_ranges(i)
        assert r["first_uoff"] == start_u
        assert r["coff_begin"] == (0 if i == 0 else want[i - 1][1])
        if end_c is not None:
            assert r["coff_end"] == end_c
    # contiguous cover of the block table
    rs = [plan.partition_ranges(i)[0] for i in range(len(want))]
    assert rs[0]["block_begin"] == 0 and all(a["block_end"] == b["block_begin"] for a, b in zip(rs, rs[1:]))


def test_fastq_without_gzi_is_sequential_and_corrupt_gzi_is_refused(tmp_path):
    import shutil
    import bamscan
    f = tmp_path / "s.fastq.bgz"
    shutil.copy(GOLDEN / "fastq" / "sample.fastq.bgz", f)
    plan = bamscan.FastqTableProvider(str(f)).scan(None, None, None, target_partitions=4, partition_mode="reference")
    assert plan.output_partition_count() == 1                            # no GZI -> FastqPartitionStrategy::Sequential
    gzi = bytearray((GOLDEN / "fastq" / "sample.fastq.bgz.gzi").read_bytes())
    gzi[8] ^= 1                                                          # first entry no longer names a member
    Path(str(f) + ".gzi").write_bytes(bytes(gzi))
    with pytest.raises(bamscan.BamScanError):
        bamscan.FastqTableProvider(str(f)).scan(None, None, None, target_partitions=4, partition_mode="reference")
    Path(str(f) + ".gzi").write_bytes(bytes(gzi[:-3]))
    with pytest.raises(bamscan.BamScanError):
        bamscan.FastqTableProvider(str(f)).scan(None, None, None, target_partitions=4, partition_mode="reference")


def _arrow_name(t: pa.DataType) -> str:
    names = {pa.utf8(): "Utf8", pa.int32(): "Int32", pa.uint32(): "UInt32", pa.float32(): "Float32", pa.int8(): "Int8", pa.uint8(): "UInt8",
             pa.int16(): "Int16", pa.uint16(): "UInt16"}
    return f"List<{_arrow_name(t.value_type)}>" if pa.types.is_list(t) else names[t]


@pytest.mark.parametrize("name,sample", [("10x_pbmc_tags.bam", 50), ("nanopore_custom_tags.bam", None), ("bam_with_tags.bam", 3), ("multi_chrom_large.bam", 100)])
def test_describe_lists_core_columns_and_every_sampled_tag(name, sample):
    """BamTableProvider::describe (table_provider.rs:703-927; tag_tests.rs:730-792): 12 core rows, then every aux tag of the first
    `sample_size` records sorted by name, typed by its first occurrence -- against the oracle's pure-python aux walk."""
    from oracle.bam_oracle import OracleBam, kind_to_arrow, load_registry
    p = _provider(GOLDEN / name)
    t = p.describe(None, sample)
    assert t.schema.names == ["column_name", "data_type", "nullable", "category", "sam_type", "description"]
    assert [f.nullable for f in t.schema] == [False, False, False, False, True, False]
    rows = t.to_pylist()
    assert [r["column_name"] for r in rows[:12]] == p.schema().names[:12]
    assert all(r["category"] == "core" and r["sam_type"] is None for r in rows[:12])
    assert [(r["data_type"], r["nullable"]) for r in rows[:12]] == [(_arrow_name(f.type), f.nullable) for f in list(p.schema())[:12]]
    assert rows[2]["description"] == "Leftmost mapping position (0-based)"
    want = {}
    o = OracleBam(str(GOLDEN / name))
    for aux in o._iter_aux(sample or 100):
        for tag, v in aux.items():
            want.setdefault(tag, v)
    reg = load_registry()
    got = rows[12:]
    assert [r["column_name"] for r in got] == sorted(want) and len(got) > 0
    for r in got:
        sam, kind = want[r["column_name"]]
        assert (r["sam_type"], r["data_type"], r["nullable"], r["category"]) == (sam, _arrow_name(kind_to_arrow(kind)), True, "tag")
        known = reg.get(r["column_name"])
        assert r["description"] == (known[2] if known else f"Custom/unknown tag ({sam})")


def test_describe_follows_provider_options_and_inferred_constructor():
    import bamscan
    p = _provider(GOLDEN / "multi_chrom.bam", zero_based=False, binary_cigar=True)
    rows = p.describe().to_pylist()
    assert rows[2]["description"] == "Leftmost mapping position (1-based)"
    assert (rows[5]["data_type"], rows[5]["description"]) == ("Binary", "CIGAR (binary LE u32 ops)")
    # try_new_with_inferred_schema (table_provider.rs:568-621) == new(.., infer_tag_types = true, sample_size, no hints)
    a = bamscan.BamTableProvider.try_new_with_inferred_schema(str(GOLDEN / "10x_pbmc_tags.bam"), None, True, ["CB", "xf", "ZZ"], None, False)
    b = _provider(GOLDEN / "10x_pbmc_tags.bam", tag_fields=["CB", "xf", "ZZ"])
    assert a.schema().equals(b.schema(), check_metadata=True)
    assert a.schema().field("xf").type == pa.int32()
    with pytest.raises(bamscan.BamScanError):
        bamscan.FastqTableProvider(str(GOLDEN / "fastq" / "sample.fastq.bgz")).describe()


def test_fastq_gzi_with_an_end_of_file_entry(tmp_path):
    """A GZI that also lists the end of the file (no member there) plans an empty last run instead of failing."""
    import shutil
    import struct
    import bamscan
    f = tmp_path / "s.fastq.bgz"
    shutil.copy(GOLDEN / "fastq" / "sample.fastq.bgz", f)
    gzi = (GOLDEN / "fastq" / "sample.fastq.bgz.gzi").read_bytes()
    n, = struct.unpack_from("<Q", gzi, 0)
    data, off, inflated = f.read_bytes(), 0, 0
    while off < len(data):                                               # BSIZE chain: total inflated size of the file
        bs = struct.unpack_from("<H", data, off + 16)[0] + 1
        inflated += struct.unpack_from("<I", data, off + bs - 4)[0]
        off += bs
    Path(str(f) + ".gzi").write_bytes(struct.pack("<Q", n + 1) + gzi[8:] + struct.pack("<QQ", len(data), inflated))
    plan = bamscan.FastqTableProvider(str(f)).scan(None, None, None, target_partitions=n + 2, partition_mode="reference")
    assert plan.output_partition_count() == n + 2
    assert plan.partition_ranges(n + 1) == []                            # the run that starts at the end of the file
    assert sum(len(plan.partition_ranges(i)) for i in range(n + 2)) == n + 1
