"""CPU: the WRITE-path oracle (oracle/bam_write_oracle.py) pinned the way the reference pins its writer
(bio-format-bam/tests/write_test.rs: write -> read round trips), plus a stronger pin the reference does not have: re-encoding
the rows of the reference's fixtures (written by htslib) reproduces the fixtures' own record bytes."""
import struct

import pyarrow as pa
import pytest

from conftest import GOLDEN, make_edge_bam
from oracle import bam_write_oracle as W
from oracle.bam_oracle import OracleBam

FIXTURES = ["multi_chrom.bam", "bam_with_tags.bam", "10x_pbmc_tags.bam", "nanopore_custom_tags.bam", "multi_chrom_large.bam", "no_coor_only.bam"]


def records_of(path):
    """-> (header_bytes, [record bytes incl. block_size]) of a BAM file, through the strict BGZF reader."""
    stream, _ = W.inflate_bgzf(open(path, "rb").read())
    l_text = struct.unpack_from("<i", stream, 4)[0]
    p = 8 + l_text
    n_ref = struct.unpack_from("<i", stream, p)[0]; p += 4
    for _ in range(n_ref):
        l = struct.unpack_from("<i", stream, p)[0]; p += 4 + l + 4
    hdr, recs = stream[:p], []
    while p < len(stream):
        bs = struct.unpack_from("<i", stream, p)[0]
        recs.append(stream[p:p + 4 + bs]); p += 4 + bs
    return hdr, recs


def core_len(rec):
    l_name, n_cig, l_seq = rec[12], struct.unpack_from("<H", rec, 16)[0], struct.unpack_from("<i", rec, 20)[0]
    return 36 + l_name + 4 * n_cig + (l_seq + 1) // 2 + l_seq


@pytest.mark.parametrize("name", FIXTURES)
@pytest.mark.parametrize("zero_based", [True, False])
def test_reencoding_fixture_rows_reproduces_the_fixture_bytes(name, zero_based):
    o = OracleBam(str(GOLDEN / name), zero_based=zero_based, tag_fields=[])
    batch = o.scan(None)
    _hdr, recs = records_of(GOLDEN / name)
    assert batch.num_rows == len(recs)
    _t, names, _l = W.build_bam_header(o.schema)
    out = W.encode_batch(batch, names, [], zero_based)
    p = 0
    for i, orig in enumerate(recs):
        bs = struct.unpack_from("<I", out, p)[0]
        mine = out[p:p + 4 + bs]; p += 4 + bs
        n = core_len(orig)
        assert bs == n - 4                                            # no tags requested: the record ends behind the qualities
        a, b = bytearray(orig[4:n]), bytearray(mine[4:])
        span = sum(struct.unpack_from("<I", orig, 36 + orig[12] + 4 * k)[0] >> 4 for k in range(struct.unpack_from("<H", orig, 16)[0])
                   if (struct.unpack_from("<I", orig, 36 + orig[12] + 4 * k)[0] & 15) in (0, 2, 3, 7, 8))
        if span == 0:                                                 # bin of zero-span reads: htslib uses end = pos + 1 (unpinned)
            a[10:12] = b[10:12]
        assert a == b, f"record {i} of {name} differs"
    assert p == len(out)


@pytest.mark.parametrize("name", FIXTURES)
def test_round_trip_with_all_tags(name, tmp_path):
    """write_test.rs::test_read_add_write_read_round_trip: read with tags, write, read again -> the same batch and schema."""
    tags = {"bam_with_tags.bam": ["NM", "MD", "AS", "XS"], "10x_pbmc_tags.bam": ["CB", "UB", "NH", "xf"],
            "nanopore_custom_tags.bam": ["pt", "NM", "de", "MM", "ML"]}.get(name, ["NM", "MD"])
    o = OracleBam(str(GOLDEN / name), tag_fields=tags)
    batch = o.scan(None)
    out = tmp_path / "out.bam"
    n = W.write_bam(out, [batch], o.schema, tags, True, {"bio.bam.sort_order": "unsorted"})
    assert n == batch.num_rows
    o2 = OracleBam(str(out), tag_fields=tags)
    back = o2.scan(None)
    assert back.num_rows == batch.num_rows
    for i, f in enumerate(batch.schema):
        assert back.column(i).equals(batch.column(i)), f.name
    assert o2.schema.metadata[b"bio.bam.sort_order"] == b"unsorted"            # write_test.rs::test_sort_on_write_false_sets_unsorted
    assert [f.type for f in o2.schema] == [f.type for f in o.schema]


def test_edge_bam_round_trip(tmp_path):
    src = tmp_path / "edge.bam"
    make_edge_bam(src, missing_qual=False)
    tags = ["NM", "MD", "XS", "XF", "XB", "XA", "XI", "XU"]
    o = OracleBam(str(src), tag_fields=tags)
    batch = o.scan(None)
    out = tmp_path / "out.bam"
    W.write_bam(out, [batch], o.schema, tags)
    back = OracleBam(str(out), tag_fields=tags).scan(None)
    assert back.equals(batch)
    # header lines survive: reference sequences, read group, program, comment
    hdr, _ = records_of(out)
    text = hdr[8:8 + struct.unpack_from("<i", hdr, 4)[0]].decode()
    assert "@SQ\tSN:chrA\tLN:1000\tM5:abc" in text and "@RG\tID:g1\tSM:s\tPL:ILLUMINA" in text and "@CO\thello\tworld" in text


def _schema(tag_fields=()):
    core = [pa.field("name", pa.string()), pa.field("chrom", pa.string()), pa.field("start", pa.uint32()), pa.field("end", pa.uint32()),
            pa.field("flags", pa.uint32(), False), pa.field("cigar", pa.string(), False), pa.field("mapping_quality", pa.uint32(), False),
            pa.field("mate_chrom", pa.string()), pa.field("mate_start", pa.uint32()), pa.field("sequence", pa.string(), False),
            pa.field("quality_scores", pa.string(), False), pa.field("template_length", pa.int32(), False)]
    md = {"bio.bam.reference_sequences": '[{"name":"chr1","length":249250621}]'}
    return pa.schema(core + list(tag_fields), metadata=md)


def _row(**kw):
    d = dict(name="read1", chrom="chr1", start=100, end=110, flags=0, cigar="10M", mapping_quality=60, mate_chrom=None, mate_start=None,
             sequence="ACGTACGTAC", quality_scores="!!!!!!!!!!", template_length=0)
    d.update(kw)
    return d


def test_flag_overflow_is_rejected():
    """serializer.rs::test_batch_to_bam_records_rejects_flag_overflow."""
    s = _schema()
    b = pa.RecordBatch.from_pylist([_row(flags=65536)], schema=s)
    with pytest.raises(W.WriteError, match="does not fit into 16-bit SAM flags"):
        W.encode_batch(b, ["chr1"], [], True)


def test_basic_record_layout():
    """serializer.rs::test_batch_to_bam_records_basic + the byte layout of SAMv1 4.2."""
    s = _schema()
    rec = W.encode_batch(pa.RecordBatch.from_pylist([_row()], schema=s), ["chr1"], [], True)
    bs, ref, pos, l_name, mapq, bin_, n_cig, flag, l_seq, nref, npos, tlen = struct.unpack_from("<IiiBBHHHiiii", rec, 0)
    assert (bs, ref, pos, l_name, mapq, n_cig, flag, l_seq, nref, npos, tlen) == (len(rec) - 4, 0, 100, 6, 60, 1, 0, 10, -1, -1, 0)
    assert bin_ == 4681 + (100 >> 14)
    assert rec[36:42] == b"read1\0" and struct.unpack_from("<I", rec, 42)[0] == (10 << 4)
    assert rec[46:51] == bytes([0x12, 0x48, 0x12, 0x48, 0x12]) and rec[51:61] == bytes(10)
    one_based = W.encode_batch(pa.RecordBatch.from_pylist([_row()], schema=s), ["chr1"], [], False)
    assert struct.unpack_from("<i", one_based, 8)[0] == 99
    # '*' / NULL name, '=' mate, unknown chromosome, '*' sequence + qualities, missing qualities of a present sequence
    r = W.encode_batch(pa.RecordBatch.from_pylist([_row(name=None, mate_chrom="=", mate_start=5, chrom="chrZ", sequence="*", quality_scores="*", cigar="*")], schema=s), ["chr1"], [], True)
    assert struct.unpack_from("<i", r, 4)[0] == -1 and r[12] == 2 and r[36:38] == b"*\0" and struct.unpack_from("<i", r, 24)[0] == -1
    r = W.encode_batch(pa.RecordBatch.from_pylist([_row(quality_scores="*")], schema=s), ["chr1"], [], True)
    assert r[51:61] == b"\xff" * 10


def test_tag_types_follow_the_field_metadata():
    """write_test.rs::test_full_tag_type_round_trip: the SAM type letter of bio.bam.tag.type decides the aux encoding."""
    def tf(name, typ, spec):
        return pa.field(name, typ, True, {"bio.bam.tag.tag": name, "bio.bam.tag.type": spec})
    fields = [tf("Xc", pa.int32(), "c"), tf("XC", pa.int32(), "C"), tf("Xs", pa.int32(), "s"), tf("XS", pa.int32(), "S"), tf("Xi", pa.int32(), "i"),
              tf("XI", pa.uint32(), "I"), tf("Xf", pa.float32(), "f"), tf("XZ", pa.string(), "Z"), tf("XH", pa.string(), "H"), tf("XA", pa.string(), "A"),
              tf("XB", pa.list_(pa.uint8()), "B:C"), tf("Xb", pa.list_(pa.int16()), "B:s"), tf("XF", pa.list_(pa.float32()), "B:f")]
    s = _schema(fields)
    row = _row(Xc=-5, XC=200, Xs=-3000, XS=60000, Xi=-70000, XI=4000000000, Xf=1.5, XZ="hello", XH="1a2b", XA="q", XB=[1, 2, 255], Xb=[-1, 300], XF=[0.5, -2.0])
    names = [f.name for f in fields]
    rec = W.encode_batch(pa.RecordBatch.from_pylist([row], schema=s), ["chr1"], names, True)
    aux = rec[61:]
    expect = (b"Xcc\xfb" + b"XCC\xc8" + b"Xss" + struct.pack("<h", -3000) + b"XSS" + struct.pack("<H", 60000) + b"Xii" + struct.pack("<i", -70000) +
              b"XII" + struct.pack("<I", 4000000000) + b"Xff" + struct.pack("<f", 1.5) + b"XZZhello\0" + b"XHH1A2B\0" + b"XAAq" +
              b"XBBC" + struct.pack("<I", 3) + bytes([1, 2, 255]) + b"XbBs" + struct.pack("<Ihh", 2, -1, 300) + b"XFBf" + struct.pack("<Iff", 2, 0.5, -2.0))
    assert aux == expect
    # NULL tags are skipped, order follows tag_fields, a value outside its SAM type fails
    rec2 = W.encode_batch(pa.RecordBatch.from_pylist([_row(XZ="z", Xc=1)], schema=s), ["chr1"], ["XZ", "Xc", "NM"], True)
    assert rec2[61:] == b"XZZz\0Xcc\x01"
    with pytest.raises(W.WriteError, match="does not fit SAM type 'c'"):
        W.encode_batch(pa.RecordBatch.from_pylist([_row(Xc=128)], schema=s), ["chr1"], ["Xc"], True)


def test_header_defaults_and_bgzf_blocks(tmp_path):
    s = pa.schema(_schema())
    text, names, lens = W.build_bam_header(pa.schema(list(s), metadata={}))
    assert text == "@HD\tVN:1.6\n" and names == [] and lens == []
    batch = pa.RecordBatch.from_pylist([_row(name=f"r{i}", sequence="ACGT" * 25, quality_scores="I" * 100, cigar="100M") for i in range(3000)], schema=s)
    out = tmp_path / "o.bam"
    W.write_bam(out, [batch.slice(0, 1000), batch.slice(1000)], s, [])
    stream, sizes = W.inflate_bgzf(out.read_bytes())
    assert all(x == 0xff00 for x in sizes[:-2]) and 0 < sizes[-2] <= 0xff00 and sizes[-1] == 0    # full blocks, a tail, the EOF marker
    assert OracleBam(str(out), tag_fields=[]).scan(None).num_rows == 3000


@pytest.mark.parametrize("name", FIXTURES)
def test_python_mirror_builds_the_same_header_as_the_oracle(name):
    """bamscan.build_bam_header (the host mirror of header_builder.rs the tests drive the C ABI with) against the oracle's
    restatement, on the schema of every fixture and with the insert_into overrides; defaults for an empty schema."""
    import bamscan
    o = OracleBam(str(GOLDEN / name), tag_fields=["NM"])
    for ov in (None, {"bio.bam.sort_order": "unsorted"}, {"bio.bam.sort_order": "coordinate", "bio.bam.file_format_version": "1.5"}):
        assert bamscan.build_bam_header(o.schema, ["NM"], ov) == W.build_bam_header(o.schema, ov)
    bare = pa.schema(list(o.schema))
    assert bamscan.build_bam_header(bare) == ("@HD\tVN:1.6\n", [], []) == W.build_bam_header(bare)
    bad = pa.schema(list(o.schema), metadata={"bio.bam.file_format_version": "x.y", "bio.bam.reference_sequences": "not json"})
    assert bamscan.build_bam_header(bad) == W.build_bam_header(bad) == ("@HD\tVN:1.6\n", [], [])


def test_edge_header_lines_survive_both_builders(tmp_path):
    import bamscan
    src = tmp_path / "edge.bam"
    make_edge_bam(src)
    o = OracleBam(str(src), tag_fields=[])
    text, names, lens = bamscan.build_bam_header(o.schema)
    assert (text, names, lens) == W.build_bam_header(o.schema)
    assert names == ["chrA", "chrB"] and lens == [1000, 2000]
    assert "@PG\tID:p\tPN:prog\tVN:1\tCL:cmd \"q\"" in text and text.startswith("@HD\tVN:1.6\tSO:unsorted\n")


def _aux_signature(rec):
    """[(tag, type letter, subtype)] of a record's aux fields, in file order."""
    p, out = core_len(rec), []
    while p < len(rec):
        tag, t = rec[p:p + 2].decode(), chr(rec[p + 2]); p += 3
        sub = None
        if t in "cCA": p += 1
        elif t in "sS": p += 2
        elif t in "iIf": p += 4
        elif t in "ZH": p = rec.index(b"\0", p) + 1
        elif t == "B":
            sub = chr(rec[p]); n = struct.unpack_from("<I", rec, p + 1)[0]
            p += 5 + n * {"c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4}[sub]
        out.append((tag, t, sub))
    return out


@pytest.mark.parametrize("name", ["bam_with_tags.bam", "10x_pbmc_tags.bam", "nanopore_custom_tags.bam", "multi_chrom_large.bam"])
def test_reencoding_with_the_files_own_type_letters_reproduces_whole_records(name):
    """Aux bytes pinned against a third-party writer too: with every tag of the first record requested in file order, and the
    columns' bio.bam.tag.type set to the letters the file uses, every record whose aux fields have that same signature must come
    back byte for byte -- block_size, core fields AND aux."""
    _hdr, recs = records_of(GOLDEN / name)
    sig = next(s for s in map(_aux_signature, recs) if s)
    tags = [t for t, _l, _s in sig]
    assert len(set(tags)) == len(tags)
    o = OracleBam(str(GOLDEN / name), tag_fields=tags)
    batch = o.scan(None)
    fields = list(o.schema)[:12]
    for tag, letter, sub in sig:
        f = o.schema.field(tag)
        spec = f"B:{sub}" if letter == "B" else letter
        fields.append(pa.field(tag, f.type, True, {"bio.bam.tag.tag": tag, "bio.bam.tag.type": spec}))
    schema = pa.schema(fields, metadata=o.schema.metadata)
    b2 = pa.RecordBatch.from_arrays([batch.column(batch.schema.get_field_index(f.name)) for f in fields], schema=schema)
    keep = [_aux_signature(r) == sig for r in recs]                      # (htslib picks the narrowest integer letter per VALUE)
    b2 = b2.filter(pa.array(keep))
    _t, names, _l = W.build_bam_header(schema)
    out = W.encode_batch(b2, names, tags, True)
    p = same = 0
    for orig in (r for r, k in zip(recs, keep) if k):
        bs = struct.unpack_from("<I", out, p)[0]
        mine = out[p:p + 4 + bs]; p += 4 + bs
        span = sum(struct.unpack_from("<I", orig, 36 + orig[12] + 4 * k)[0] >> 4 for k in range(struct.unpack_from("<H", orig, 16)[0])
                   if (struct.unpack_from("<I", orig, 36 + orig[12] + 4 * k)[0] & 15) in (0, 2, 3, 7, 8))
        a = bytearray(orig)
        if span == 0:
            a[14:16] = mine[14:16]                                   # bin of zero-span reads (unpinned)
        assert bytes(a) == mine, f"{name}: a record with the first record's aux signature differs"
        same += 1
    assert p == len(out) and same >= len(recs) // 20 and same > 0, (same, len(recs))
