"""GPU parity at the benchmarked shapes, through the DEFAULT product path (no chunk cap, no forced kernels).

VERDICT r1 "what's weak" 1: the 100 M-read bench ran kernels (full inflate waves, automatic kernel choice, 768 MiB slices,
carry between waves) that the fixture-sized tests only exercised in miniature.  Here the config-2 generator is scanned at a
size that gives two full inflate waves plus a tail chunk (so the lane-group AND the CTA-per-member inflate kernels both run,
as they do in the bench), and the long-read generator at > 20 k reads; the whole output is compared with the oracle
(oracle/verify.py: per-column streaming checksums over every row + full equality on three 1 M-row windows).
Reference semantics: physical_exec.rs:371-598.
"""
import os

import pyarrow as pa
import pytest

from conftest import GOLDEN, gen_bam, make_edge_bam

pytestmark = pytest.mark.gpu

TAGS = ["NM", "MD", "AS", "RG"]


def _verify(path, tags, n_parts, **kw):
    import bamscan
    from oracle.bam_oracle import OracleBam
    from oracle.verify import verify_full_scan
    o = OracleBam(str(path), tag_fields=tags)
    p = bamscan.BamTableProvider(str(path), None, True, tags, False, True, 100, None, index_path="", **kw)
    plan = p.scan(None, [], None)
    rep = verify_full_scan(o, str(path) + ".bai", plan.execute(0), n_parts=n_parts, threads=os.cpu_count())
    st = plan.last_stats
    p.close()
    return rep, st


def test_default_path_two_inflate_waves_short_reads(syn_dir):
    reads = int(os.environ.get("BAMSCAN_SCALE_READS", 9_000_000))      # 22 496 members per wave ~ 4.28 M reads
    path = gen_bam(syn_dir, "short", reads, seed=2, bai=True)
    rep, st = _verify(path, TAGS, n_parts=max(8, reads // 1_000_000))
    assert rep["rows"] == reads
    if reads >= 9_000_000:
        assert st["chunks"] >= 3, st          # two whole waves + the tail chunk
    print("short reads:", rep["rows"], "rows,", st["chunks"], "chunks,", st["batches"], "batches; windows", rep["windows_compared"])


def test_default_path_long_reads(syn_dir):
    reads = int(os.environ.get("BAMSCAN_SCALE_LONG_READS", 24_000))
    path = gen_bam(syn_dir, "long", reads, seed=5, bai=True)
    rep, st = _verify(path, ["NM", "MD", "MM", "ML"], n_parts=8)
    assert rep["rows"] == reads
    print("long reads:", rep["rows"], "rows,", st["chunks"], "chunks; windows", rep["windows_compared"])


@pytest.mark.parametrize("inflate", [16, 4, 8], ids=["cta_per_member", "warp_per_member", "lane_group"])
@pytest.mark.parametrize("fixed", [0, 2], ids=["thread_per_record", "warp_per_record"])
@pytest.mark.parametrize("zero_based", [True, False])
def test_edge_bam_on_gpu(tmp_path, inflate, fixed, zero_based):
    """The hand-made edge-case BAM (MAPQ 255, name "*", l_seq 0, 268435455M, negative c, B:S, I = 4e9, CIGAR-less placed
    read, 0xFF-filled qualities) under every inflate kernel and both decode_fixed kernels (sam_read_test.rs:825-830,
    902-907, 988-990)."""
    import bamscan
    from oracle.bam_oracle import OracleBam, schema_equal
    path = tmp_path / "edge.bam"
    make_edge_bam(path, missing_qual=True)
    tags = ["NM", "MD", "XS", "XF", "XB", "XA", "XI", "XU", "ZZ"]
    o = OracleBam(str(path), zero_based=zero_based, tag_fields=tags)
    want = pa.Table.from_batches([o.scan()])
    p = bamscan.BamTableProvider(str(path), None, zero_based, tags, False, True, 100, None, index_path="", debug_flags=inflate | fixed)
    assert schema_equal(p.schema(), o.schema)
    got = p.scan(None, [], None).collect()
    assert got.num_rows == want.num_rows == 6
    for name in want.schema.names:
        assert got[name].combine_chunks().equals(want[name].combine_chunks()), f"{name}: {got[name].to_pylist()} != {want[name].to_pylist()}"
    assert got["quality_scores"].to_pylist()[4] == " " * 40 and got["mapping_quality"].to_pylist()[0] == 255 and got["name"].to_pylist()[1] == "*"
    p.close()


def test_float_tags_into_utf8_columns(tmp_path):
    """sam_tag_io.rs:670-676: a float tag read into a Utf8 column (unknown tag, inference off) is Rust's f32::to_string():
    shortest digits that round-trip, no exponent.  GPU (f32_text.h) == oracle (rust_f32_to_string)."""
    import bamscan
    from oracle.bam_oracle import OracleBam
    path = tmp_path / "edge.bam"
    make_edge_bam(path, missing_qual=True)
    tags = ["XF"] + ["Y" + chr(ord("a") + i) for i in range(10)]
    o = OracleBam(str(path), tag_fields=tags, infer_tag_types=False)
    want = pa.Table.from_batches([o.scan()])
    p = bamscan.BamTableProvider(str(path), None, True, tags, False, False, 100, None, index_path="")
    got = p.scan(None, [], None).collect()
    for name in tags:
        assert want.schema.field(name).type == pa.string()
        assert got[name].combine_chunks().equals(want[name].combine_chunks()), f"{name}: {got[name].to_pylist()} != {want[name].to_pylist()}"
    assert got["XF"].to_pylist()[0] == "1.5" and got["Ya"].to_pylist()[5] == "0.0000001" and got["Yc"].to_pylist()[5] == "0.1"
    p.close()


def test_decode_all_tag_fields_switch(tmp_path):
    """sam_tag_io.rs:42-52: once any tag column is projected the reference decodes EVERY tag_fields entry, so a value that
    does not fit an unprojected tag column's type fails the scan.  Default here: only projected tags are decoded (the scan
    succeeds); decode_all_tag_fields=True reproduces the reference (and the oracle): the scan fails."""
    import bamscan
    from oracle.bam_oracle import OracleBam
    path = tmp_path / "edge.bam"
    make_edge_bam(path, missing_qual=True)
    tags, hints = ["NM", "XU"], ["XU:i"]          # XU is an 'I' of 4e9: does not fit Int32
    proj = [0, 12]                                 # name + NM; XU is in tag_fields but not projected
    o = OracleBam(str(path), tag_fields=tags, infer_tag_types=False, tag_type_hints=hints)
    with pytest.raises(Exception):
        o.scan(proj)
    p = bamscan.BamTableProvider(str(path), None, True, tags, False, False, 100, hints, index_path="")
    got = p.scan(proj, [], None).collect()
    assert got.num_rows == 6 and got.column_names == ["name", "NM"]
    p.close()
    p = bamscan.BamTableProvider(str(path), None, True, tags, False, False, 100, hints, index_path="", decode_all_tag_fields=True)
    with pytest.raises(bamscan.BamScanError):
        p.scan(proj, [], None).collect()
    p.close()
    # a projection without any tag column decodes no tag at all, in the reference too
    p = bamscan.BamTableProvider(str(path), None, True, tags, False, False, 100, hints, index_path="", decode_all_tag_fields=True)
    assert p.scan([0, 2], [], None).collect().num_rows == 6
    assert OracleBam(str(path), tag_fields=tags, infer_tag_types=False, tag_type_hints=hints).scan([0, 2]).num_rows == 6
    p.close()


def test_twenty_tag_columns():
    """Round 1 stopped at 16 projected tag columns; the limit is 32 now (present and absent tags, all column kinds)."""
    import bamscan
    from oracle.bam_oracle import OracleBam, schema_equal
    path = GOLDEN / "10x_pbmc_tags.bam"
    tags = ["RE", "ts", "CB", "NH", "xf", "NM", "MD", "AS", "RG", "CR", "CY", "UB", "UR", "UY", "GX", "GN", "fx", "HI", "nM", "TX", "AN", "pa"]
    o = OracleBam(str(path), tag_fields=tags)
    want = pa.Table.from_batches([o.scan()])
    p = bamscan.BamTableProvider(str(path), None, True, tags, False, True, 100, None, index_path="")
    assert schema_equal(p.schema(), o.schema)
    got = p.scan(None, [], None).collect()
    assert got.num_rows == want.num_rows and len(got.schema) == 12 + len(tags)
    for name in want.schema.names:
        assert got[name].combine_chunks().equals(want[name].combine_chunks()), name
    p.close()
