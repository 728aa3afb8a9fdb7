"""CPU: pins the oracle (oracle/bam_oracle.c) against the committed golden facts of the reference's own fixtures
(tests/golden/golden_facts.json, produced by the independent pure-python restatement tests/golden/make_golden.py)
and against the value pins of the reference's tests."""
import collections
import hashlib
import json
import struct
import zlib

import pyarrow as pa
import pyarrow.compute as pc
import pytest

from conftest import GOLDEN

FACTS = json.loads((GOLDEN / "golden_facts.json").read_text())


def _oracle(path, **kw):
    from oracle.bam_oracle import OracleBam
    return OracleBam(str(path), **kw)


def _sha(s):
    return hashlib.sha256(s.encode()).hexdigest()[:16]


@pytest.mark.parametrize("name", sorted(FACTS))
def test_oracle_matches_golden_facts(name):
    f = FACTS[name]
    assert hashlib.sha256((GOLDEN / name).read_bytes()).hexdigest()[:16] == f["sha256_16"]
    t = pa.Table.from_batches([_oracle(GOLDEN / name, zero_based=True).scan()])
    assert t.num_rows == f["records"]
    counts = collections.Counter(str(c) for c in t["chrom"].to_pylist())
    assert dict(counts) == f["per_chrom"]
    s = lambda c: pc.sum(pc.cast(t[c], pa.int64())).as_py() or 0
    assert s("start") == f["sum_start"] and s("end") == f["sum_end"]
    assert t["end"].null_count == f["null_end"] and t["start"].null_count == f["null_start"]
    assert s("flags") == f["sum_flags"] and s("mapping_quality") == f["sum_mapq"]
    assert s("mate_start") == f["sum_mate_start"] and s("template_length") == f["sum_tlen"]
    assert sum(len(x) for x in t["name"].to_pylist()) == f["bytes_name"]
    assert _sha("".join(t["sequence"].to_pylist())) == f["sha_seq"]
    assert _sha("".join(t["quality_scores"].to_pylist())) == f["sha_qual"]
    assert _sha(",".join(t["cigar"].to_pylist())) == f["sha_cigar"]
    assert _sha(",".join(t["name"].to_pylist())) == f["sha_name"]
    first = t.slice(0, 1).to_pylist()[0]
    for k_or, k_f in (("name", "name"), ("chrom", "chrom"), ("start", "start"), ("end", "end"), ("flags", "flags"), ("cigar", "cigar"),
                      ("mapping_quality", "mapq"), ("mate_chrom", "mate_chrom"), ("mate_start", "mate_start"), ("template_length", "tlen")):
        assert first[k_or] == f["first"][k_f], k_or
    last = t.slice(t.num_rows - 1, 1).to_pylist()[0]
    assert last["name"] == f["last"]["name"] and last["cigar"] == f["last"]["cigar"] and last["template_length"] == f["last"]["tlen"]


def test_reference_row_count_pins():
    # indexed_read_test.rs:76-77,108,121 ; indexed_read_large_test.rs:63,85,95 ; tag_tests.rs row counts
    assert FACTS["multi_chrom.bam"]["records"] == 421
    assert FACTS["multi_chrom.bam"]["per_chrom"] == {"chr1": 160, "chr2": 159, "chrX": 102}
    assert FACTS["multi_chrom_large.bam"]["records"] == 4277
    assert FACTS["multi_chrom_large.bam"]["per_chrom"] == {"chr1": 1662, "chr2": 1694, "chrX": 921}
    assert [FACTS[n]["records"] for n in ("nanopore_custom_tags.bam", "bam_with_tags.bam", "10x_pbmc_tags.bam", "no_coor_only.bam")] == [20, 14, 10, 2]


def test_one_based_coordinates_shift_start_but_not_end():
    # physical_exec.rs:428-463: start follows the coordinate system, end is always the 1-based inclusive end
    z = pa.Table.from_batches([_oracle(GOLDEN / "multi_chrom.bam", zero_based=True).scan(projection=[2, 3, 8])])
    o = pa.Table.from_batches([_oracle(GOLDEN / "multi_chrom.bam", zero_based=False).scan(projection=[2, 3, 8])])
    assert pc.sum(pc.subtract(pc.cast(o["start"], pa.int64()), pc.cast(z["start"], pa.int64()))).as_py() == 421
    assert o["end"].equals(z["end"])
    assert pc.sum(pc.subtract(pc.cast(o["mate_start"], pa.int64()), pc.cast(z["mate_start"], pa.int64()))).as_py() == 421 - z["mate_start"].null_count


def test_schema_sizes_and_tag_types():
    # tag_tests.rs: 12 core fields (+1 per tag), registry types, inferred types for unknown tags
    from oracle.bam_oracle import K_Int32
    assert len(_oracle(GOLDEN / "10x_pbmc_tags.bam").schema) == 12
    o = _oracle(GOLDEN / "10x_pbmc_tags.bam", tag_fields=["NM", "MD"])
    assert len(o.schema) == 14 and o.schema.field("NM").type == pa.int32() and o.schema.field("MD").type == pa.utf8()
    assert o.schema.field("NM").metadata[b"bio.bam.tag.type"] == b"i"
    assert o.schema.field("NM").metadata[b"bio.bam.tag.description"] == b"Edit distance to the reference"
    o = _oracle(GOLDEN / "nanopore_custom_tags.bam", tag_fields=["pa", "ML", "FZ", "zz"])
    assert o.schema.field("pa").type == pa.list_(pa.field("item", pa.int32(), True))          # inferred from the file (B:i)
    assert o.schema.field("pa").metadata[b"bio.bam.tag.type"] == b"B:i"
    assert o.schema.field("ML").type == pa.list_(pa.field("item", pa.uint8(), True))          # registry
    assert o.schema.field("FZ").type == pa.list_(pa.field("item", pa.uint16(), True))
    assert o.schema.field("zz").type == pa.utf8() and o.schema.field("zz").metadata[b"bio.bam.tag.description"] == b"Unknown tag"
    md = o.schema.metadata
    assert md[b"bio.coordinate_system_zero_based"] == b"true" and b"bio.bam.reference_sequences" in md
    refs = json.loads(md[b"bio.bam.reference_sequences"])
    assert len(refs) == 19 and set(refs[0]) >= {"name", "length"}


def test_hint_grammar():
    # tag_registry.rs:698-752 and its unit tests (:820-909)
    from oracle.bam_oracle import parse_tag_type_hints, K_Int32, K_UInt32, K_Float32, K_Utf8, K_ListUInt8
    h = parse_tag_type_hints(["pt:i", "de:f", "XA:Z", "ml:B:C", "xs:A", "ui:I"])
    assert h == {"pt": ("i", K_Int32), "de": ("f", K_Float32), "XA": ("Z", K_Utf8), "ml": ("B", K_ListUInt8), "xs": ("A", K_Utf8), "ui": ("I", K_UInt32)}
    for bad in ["pt", "pt:ii", "pt:B", "pt:B:x", "pt:q", "a:b:c:d"]:
        with pytest.raises(ValueError):
            parse_tag_type_hints([bad])


from conftest import make_edge_bam   # hand-made BAM with the edge cases the reference pins (also run through the GPU: test_gpu_scale.py)


def test_edge_case_values(tmp_path):
    p = tmp_path / "edge.bam"
    make_edge_bam(p)
    o = _oracle(p, zero_based=True, tag_fields=["NM", "MD", "XS", "XF", "XB", "XA", "XI", "XU", "ZZ"])
    t = pa.Table.from_batches([o.scan()])
    rows = t.to_pylist()
    r = rows[0]
    assert (r["name"], r["chrom"], r["start"], r["end"], r["mapping_quality"], r["template_length"]) == ("r1", "chrA", 99, 99 + 19, 255, -160)   # MAPQ 255 kept (sam_read_test.rs:825-830), TLEN sign (:988-990)
    assert r["cigar"] == "5S10M2I3D4N1=1X" and r["sequence"] == "ACGTNACGTNACGTNACGT" and r["quality_scores"] == "".join(chr(33 + i) for i in range(19))
    assert (r["NM"], r["MD"], r["XS"], r["XF"], r["XB"], r["XA"], r["XI"], r["XU"], r["ZZ"]) == (5, "10A5", -2, 1.5, [7, 65535], "q", -70000, 4000000000, None)
    assert t.schema.field("XB").type == pa.list_(pa.field("item", pa.uint16(), True)) and t.schema.field("XU").type == pa.uint32()
    assert (rows[1]["name"], rows[1]["chrom"], rows[1]["start"], rows[1]["end"], rows[1]["sequence"], rows[1]["quality_scores"], rows[1]["cigar"]) == ("*", None, None, None, "", "", "")
    assert rows[1]["NM"] is None
    assert rows[2]["cigar"] == "268435455M" and rows[2]["end"] == 268435455 and rows[2]["sequence"] == "A"
    assert rows[3]["start"] == 5 and rows[3]["end"] is None
    md = {k.decode(): v.decode() for k, v in t.schema.metadata.items()}
    assert md["bio.bam.file_format_version"] == "1.6" and md["bio.bam.sort_order"] == "unsorted"
    assert json.loads(md["bio.bam.reference_sequences"]) == [{"name": "chrA", "length": 1000, "other_fields": {"M5": "abc"}}, {"name": "chrB", "length": 2000}]
    assert json.loads(md["bio.bam.read_groups"]) == [{"id": "g1", "sample": "s", "platform": "ILLUMINA", "other_fields": {"XX": "y"}}]
    assert json.loads(md["bio.bam.program_info"]) == [{"id": "p", "name": "prog", "version": "1", "command_line": "cmd \"q\""}]
    assert json.loads(md["bio.bam.comments"]) == ["hello\tworld"]
    # RecordBuf-style end (the other plausible value for the unpinned case) behind the switch
    o2 = _oracle(p, end_zero_span_mode=1)
    assert o2.scan(projection=[3]).column(0).to_pylist()[3] == 5


def test_partition_ownership_rule(tmp_path):
    """Block-range partitions seeded with exact record starts reproduce the sequential scan."""
    path = GOLDEN / "multi_chrom_large.bam"
    o = _oracle(path)
    full = o.scan()
    voff, idx, nrec = o.index_records(stride=0)
    assert nrec == 4277
    cuts = [0, int(voff[len(voff) // 3]), int(voff[2 * len(voff) // 3]), 0]
    parts = [o.scan(start_voffset=cuts[i], stop_voffset=cuts[i + 1]) for i in range(3)]
    assert sum(p.num_rows for p in parts) == 4277
    assert pa.Table.from_batches(parts).equals(pa.Table.from_batches([full]))
