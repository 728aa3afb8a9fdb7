"""CPU oracle of the BAM WRITE path -- TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's
cpu_baseline leg); nothing under datafusion-bio-formats_b200/ may import it.

Restates, in plain Python:
  * batch_to_alignment_records / build_single_record   datafusion/bio-format-core/src/sam_record_serializer.rs:15-212
  * build_tag_data / arrow_to_sam_tag_value & co       datafusion/bio-format-core/src/sam_tag_io.rs:109-147, 206-520
  * build_bam_header                                   datafusion/bio-format-bam/src/header_builder.rs:43-186
  * write_bam_stream                                   datafusion/bio-format-bam/src/write_exec.rs:278-350
and, for what the reference delegates to noodles (noodles-bam 0.92.0 record encoder, noodles-sam 0.87.0 header writer,
noodles-bgzf 0.49.0 writer; git fork @42a3c016, not vendored), the published formats they implement: SAMv1 4.2 (BAM record
layout, reg2bin, 4-bit bases, 0xFF for missing qualities), 1.3 (header lines), 4.1 (BGZF, 0xff00-byte blocks + EOF marker).

PARITY PINNING: the reference holds no golden BAM bytes for this path; its tests (bio-format-bam/tests/write_test.rs) are
write -> read round trips.  This oracle is pinned the same way -- tests/test_write_oracle.py writes the reference's fixtures and
the write_test.rs vectors with it and reads them back with the (pinned) read oracle -- and, beyond that, against the BYTES of the
reference's fixtures (written by htslib): re-encoding their rows reproduces every record's core fields, and whole records (aux
included) wherever the requested type letters are the file's own.  Details only noodles decides and no
reference test fixes are listed in DESIGN.md 3.6 ("unpinned": bin of zero-span reads, base case folding, HashMap order of
extra header fields).
"""
from __future__ import annotations

import json
import struct
import zlib

import pyarrow as pa

BASES = "=ACMGRSVTWYHKDBN"
_BASE_CODE = [15] * 256
for _i, _c in enumerate(BASES):
    _BASE_CODE[ord(_c)] = _i
    _BASE_CODE[ord(_c.lower())] = _i          # SAMv1 4.2.3: case-insensitive, everything else -> N
CIGAR_OPS = "MIDNSHP=X"
UNMAPPED_BIN = 4680
BGZF_BLOCK = 0xff00                            # noodles-bgzf MAX_BUF_SIZE
BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


class WriteError(Exception):
    """DataFusionError::Execution of the reference write path."""


# ---------------------------------------------------------------------------------------------------------------------
# header (header_builder.rs:43-186 + the noodles-sam header writer)

_SQ_TAGS = {"AH", "AN", "AS", "DS", "M5", "SP", "TP", "UR"}
_RG_TAGS = {"BC", "CN", "DT", "FO", "KS", "PG", "PI", "PM", "PU"}
_PG_TAGS = {"PP", "DS"}


def _json(md, key):
    try:
        return json.loads(md[key]) if key in md else None
    except Exception:
        return None                              # from_json_string -> None on failure


def build_bam_header(schema: pa.Schema, overrides: dict | None = None):
    """-> (sam_text, ref_names, ref_lengths).  Extra per-line fields come from a HashMap in the reference (arbitrary order);
    here they are written in sorted key order."""
    md = {(k.decode() if isinstance(k, bytes) else k): (v.decode() if isinstance(v, bytes) else v) for k, v in (schema.metadata or {}).items()}
    md.update(overrides or {})
    ver = md.get("bio.bam.file_format_version", "1.6")
    parts = ver.split(".")
    if not (len(parts) == 2 and all(p.isdigit() for p in parts)):
        ver = "1.6"
    else:
        ver = f"{int(parts[0])}.{int(parts[1])}"
    hd = [f"VN:{ver}"]
    for key, tag in (("bio.bam.sort_order", "SO"), ("bio.bam.group_order", "GO"), ("bio.bam.subsort_order", "SS")):
        if key in md:
            hd.append(f"{tag}:{md[key]}")
    lines = ["@HD\t" + "\t".join(hd)]
    names, lens = [], []
    for sq in _json(md, "bio.bam.reference_sequences") or []:
        if int(sq["length"]) == 0:
            raise WriteError("Reference sequence length cannot be zero")
        f = [f"SN:{sq['name']}", f"LN:{int(sq['length'])}"]
        f += [f"{k}:{v}" for k, v in sorted((sq.get("other_fields") or {}).items()) if k in _SQ_TAGS]
        lines.append("@SQ\t" + "\t".join(f))
        names.append(sq["name"]); lens.append(int(sq["length"]))
    for rg in _json(md, "bio.bam.read_groups") or []:
        f = [f"ID:{rg['id']}"]
        for key, tag in (("sample", "SM"), ("platform", "PL"), ("library", "LB"), ("description", "DS")):
            if rg.get(key) is not None:
                f.append(f"{tag}:{rg[key]}")
        f += [f"{k}:{v}" for k, v in sorted((rg.get("other_fields") or {}).items()) if k in _RG_TAGS]
        lines.append("@RG\t" + "\t".join(f))
    for pg in _json(md, "bio.bam.program_info") or []:
        f = [f"ID:{pg['id']}"]
        for key, tag in (("name", "PN"), ("version", "VN"), ("command_line", "CL")):
            if pg.get(key) is not None:
                f.append(f"{tag}:{pg[key]}")
        f += [f"{k}:{v}" for k, v in sorted((pg.get("other_fields") or {}).items()) if k in _PG_TAGS]
        lines.append("@PG\t" + "\t".join(f))
    for co in _json(md, "bio.bam.comments") or []:
        lines.append(f"@CO\t{co}")
    return "\n".join(lines) + "\n", names, lens


def bam_header_bytes(text: str, ref_names, ref_lens) -> bytes:
    t = text.encode()
    out = [b"BAM\x01", struct.pack("<i", len(t)), t, struct.pack("<i", len(ref_names))]
    for n, l in zip(ref_names, ref_lens):
        nb = n.encode() + b"\x00"
        out += [struct.pack("<i", len(nb)), nb, struct.pack("<i", l)]
    return b"".join(out)


# ---------------------------------------------------------------------------------------------------------------------
# records (sam_record_serializer.rs:96-212 + BAM record layout)

def region_to_bin(start1: int, end1: int) -> int:
    """noodles-csi binning_index region_to_bin over 1-based closed [start, end] (SAMv1 5.3 reg2bin on start-1, end)."""
    s, e = start1 - 1, end1 - 1
    if s >> 14 == e >> 14: return ((1 << 15) - 1) // 7 + (s >> 14)
    if s >> 17 == e >> 17: return ((1 << 12) - 1) // 7 + (s >> 17)
    if s >> 20 == e >> 20: return ((1 << 9) - 1) // 7 + (s >> 20)
    if s >> 23 == e >> 23: return ((1 << 6) - 1) // 7 + (s >> 23)
    if s >> 26 == e >> 26: return ((1 << 3) - 1) // 7 + (s >> 26)
    return 0


def parse_cigar(text: str):
    """sam_record_serializer.rs:223-235 ("*" / "" -> no ops; else noodles sam::record::Cigar)."""
    if text in ("*", ""):
        return []
    ops, n, have = [], 0, False
    for ch in text:
        if ch.isdigit() and ch.isascii():
            n = n * 10 + ord(ch) - 48; have = True
        else:
            k = CIGAR_OPS.find(ch)
            if k < 0 or not have:
                raise WriteError(f"Failed to parse CIGAR '{text}'")
            if n >= 1 << 28:
                raise WriteError(f"CIGAR op length {n} does not fit 28 bits")
            ops.append((n << 4) | k); n = 0; have = False
    if have:
        raise WriteError(f"Failed to parse CIGAR '{text}'")
    return ops


_INT_RANGE = {"c": (-128, 127), "C": (0, 255), "s": (-32768, 32767), "S": (0, 65535), "i": (-2**31, 2**31 - 1), "I": (0, 2**32 - 1)}
_INT_FMT = {"c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I"}
_ARROW_SUBTYPE = {pa.int8(): "c", pa.uint8(): "C", pa.int16(): "s", pa.uint16(): "S", pa.int32(): "i", pa.uint32(): "I", pa.float32(): "f"}


def _is_int(t): return pa.types.is_integer(t)


def parse_sam_tag_type(spec: str):
    parts = spec.split(":")
    if len(parts) == 1 and len(parts[0]) == 1:
        return parts[0], None
    if len(parts) == 2 and parts[0] == "B" and len(parts[1]) == 1 and parts[1] in "cCsSiIf":
        return "B", parts[1]
    raise WriteError(f"Invalid SAM tag type metadata: '{spec}'")


def encode_tag(tag: str, value, arrow_type: pa.DataType, spec: str) -> bytes:
    """arrow_to_sam_tag_value (sam_tag_io.rs:206-235) + the BAM aux field layout; b"" when the reference inserts nothing."""
    t, sub = parse_sam_tag_type(spec)
    tb = tag.encode()
    if t in "icsCSI":
        if not _is_int(arrow_type):
            raise WriteError(f"Tag value type mismatch for integer: {arrow_type}")
        lo, hi = _INT_RANGE[t]
        if not lo <= value <= hi:
            raise WriteError(f"Integer value {value} does not fit SAM type '{t}'")
        return tb + t.encode() + struct.pack(_INT_FMT[t], value)
    if t == "f":
        if not pa.types.is_floating(arrow_type):
            raise WriteError(f"Tag value type mismatch for float: {arrow_type}")
        import math
        if arrow_type == pa.float64() and (not math.isfinite(value) or abs(value) > 3.4028234663852886e38):
            raise WriteError(f"Float value {value} does not fit SAM type 'f'")
        return tb + b"f" + struct.pack("<f", value)
    if t == "Z":
        if not pa.types.is_string(arrow_type):
            raise WriteError(f"Tag value type mismatch for string: {arrow_type}")
        return tb + b"Z" + value.encode() + b"\x00"
    if t == "H":
        if not pa.types.is_string(arrow_type):
            raise WriteError(f"Tag value type mismatch for hex string: {arrow_type}")
        v = value.upper()
        if len(v) % 2 or any(c not in "0123456789ABCDEF" for c in v):
            raise WriteError(f"Invalid SAM hex tag value '{v}'")
        return tb + b"H" + v.encode() + b"\x00"
    if t == "A":
        if pa.types.is_string(arrow_type):
            b = value.encode()
            if len(b) != 1 or b[0] >= 128:
                raise WriteError(f"Character tags must be a single ASCII byte, got '{value}'")
            return tb + b"A" + b
        if _is_int(arrow_type):
            if not 0 <= value <= 255:
                raise WriteError(f"Character tag value {value} does not fit into a single byte")
            return tb + b"A" + bytes([value])
        raise WriteError(f"Tag value type mismatch for character: {arrow_type}")
    if t == "B":
        if not pa.types.is_list(arrow_type):
            raise WriteError(f"Tag value type mismatch for array: {arrow_type}")
        st = sub or _ARROW_SUBTYPE.get(arrow_type.value_type)
        if st is None:
            raise WriteError(f"Unable to determine SAM array subtype for Arrow type {arrow_type.value_type}")
        if any(v is None for v in value):
            raise WriteError("SAM array tags cannot contain null elements")
        if st == "f":
            body = b"".join(struct.pack("<f", v) for v in value)
        else:
            lo, hi = _INT_RANGE[st]
            for v in value:
                if not _is_int(arrow_type.value_type):
                    raise WriteError(f"Unsupported array element type for SAM subtype '{st}': {arrow_type.value_type}")
                if not lo <= v <= hi:
                    raise WriteError(f"Array element {v} does not fit SAM subtype '{st}'")
            body = b"".join(struct.pack(_INT_FMT[st], v) for v in value)
        return tb + b"B" + st.encode() + struct.pack("<I", len(value)) + body
    if pa.types.is_string(arrow_type):              # unknown type letter: strings pass as Z, everything else is dropped
        return tb + b"Z" + value.encode() + b"\x00"
    return b""


def tag_columns(schema: pa.Schema, tag_fields):
    """build_tag_column_map (sam_record_serializer.rs:281-298): tag_fields entries that name a column carrying bio.bam.tag.tag."""
    out = []
    for name in tag_fields:
        idx = schema.get_field_index(name)
        if idx < 0:
            continue
        md = schema.field(idx).metadata or {}
        if b"bio.bam.tag.tag" not in md or len(name.encode()) != 2:
            continue
        out.append((name, idx, (md.get(b"bio.bam.tag.type") or b"Z").decode()))
    return out


def encode_record(row: dict, ref_map: dict, tags, zero_based: bool) -> bytes:
    """One BAM record (with its block_size prefix) from a row of python values keyed by column name; `tags` is a list of
    (name, value_or_None, arrow_type, type_spec)."""
    # `.value(row)` on a null slot of a non-nullable column reads the slot's bytes: 0 / "" as Arrow builders leave them
    row = dict(row)
    for k in ("flags", "mapping_quality", "template_length"):
        if row[k] is None: row[k] = 0
    for k in ("cigar", "sequence", "quality_scores"):
        if row[k] is None: row[k] = ""
    name = row["name"]
    name_b = b"*" if name is None or name == "*" else name.encode()
    if len(name_b) > 254:
        raise WriteError("read name longer than 254 bytes")
    flags = row["flags"]
    if flags > 0xffff:
        raise WriteError(f"Flag value {flags} does not fit into 16-bit SAM flags")
    ref_id = -1 if row["chrom"] is None else ref_map.get(row["chrom"], -1)

    def pos1(v):
        if v is None:
            return None
        p = v + 1 if zero_based else v
        if p - 1 > 0x7fffffff:
            raise WriteError("position does not fit the 32-bit BAM field")      # the encoder's i32::try_from(position - 1)
        return p if p >= 1 else None              # Position::try_from(0) fails -> None

    start = pos1(row["start"])
    mapq = row["mapping_quality"] & 0xff          # `as u8`; 255 = missing, written as 255 again
    cigar = row["cigar"]
    if isinstance(cigar, (bytes, bytearray)):
        if len(cigar) % 4:
            raise WriteError("Failed to decode binary CIGAR: length is not a multiple of 4")
        ops = list(struct.unpack(f"<{len(cigar) // 4}I", cigar))
        if any((o & 15) > 8 for o in ops):
            raise WriteError("Failed to decode binary CIGAR: invalid op")
    else:
        ops = parse_cigar(cigar)
    if len(ops) > 0xffff:
        raise WriteError("more than 65535 CIGAR ops (CG-tag overflow) is not supported by this build")
    mc = row["mate_chrom"]
    mate_ref = -1 if mc is None else (ref_id if mc == "=" else ref_map.get(mc, -1))
    mate_start = pos1(row["mate_start"])
    seq = row["sequence"]
    seq_b = b"" if seq in ("*", "") else seq.encode()
    qual = row["quality_scores"]
    qual_b = b"" if qual in ("*", "") else bytes(max(0, c - 33) for c in qual.encode())
    if qual_b and len(qual_b) != len(seq_b):
        raise WriteError("sequence-quality scores length mismatch")
    if not qual_b:
        qual_b = b"\xff" * len(seq_b)
    span = sum(o >> 4 for o in ops if (o & 15) in (0, 2, 3, 7, 8))
    bin_ = UNMAPPED_BIN
    if start is not None:
        end = start + span - 1                    # RecordBuf::alignment_end, unconditional; Position::new(0) -> None
        if end >= 1:
            bin_ = region_to_bin(start, end)
    packed = bytearray((len(seq_b) + 1) // 2)
    for i, c in enumerate(seq_b):
        packed[i >> 1] |= _BASE_CODE[c] << (0 if i & 1 else 4)
    aux = b"".join(encode_tag(n, v, t, s) for n, v, t, s in tags if v is not None)
    body = struct.pack("<iiBBHHHiiii", ref_id, -1 if start is None else start - 1, len(name_b) + 1, mapq, bin_, len(ops), flags,
                       len(seq_b), mate_ref, -1 if mate_start is None else mate_start - 1, row["template_length"])
    body += name_b + b"\x00" + struct.pack(f"<{len(ops)}I", *ops) + bytes(packed) + qual_b + aux
    return struct.pack("<I", len(body)) + body


REQUIRED = ("name", "chrom", "start", "flags", "cigar", "mapping_quality", "mate_chrom", "mate_start", "sequence", "quality_scores", "template_length")


def encode_batch(batch: pa.RecordBatch, ref_names, tag_fields, zero_based: bool) -> bytes:
    """batch_to_alignment_records + the record encoder: the concatenated BAM records of one batch."""
    if batch.num_rows == 0:
        return b""
    for c in REQUIRED:
        if batch.schema.get_field_index(c) < 0:
            raise WriteError(f"Required column '{c}' not found in batch")
    ref_map = {}
    for i, n in enumerate(ref_names):
        ref_map[n] = i                              # HashMap collect: a later duplicate wins
    cols = {c: batch.column(batch.schema.get_field_index(c)).to_pylist() for c in REQUIRED}
    tcols = [(n, batch.column(i).to_pylist(), batch.schema.field(i).type, spec) for n, i, spec in tag_columns(batch.schema, tag_fields)]
    out = []
    for r in range(batch.num_rows):
        row = {c: cols[c][r] for c in REQUIRED}
        out.append(encode_record(row, ref_map, [(n, v[r], t, s) for n, v, t, s in tcols], zero_based))
    return b"".join(out)


# ---------------------------------------------------------------------------------------------------------------------
# file (write_exec.rs:278-350 + BGZF framing)

def bgzf_member(payload: bytes, level: int = 6) -> bytes:
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    data = co.compress(payload) + co.flush()
    if len(data) + 26 > 65536:
        data = b"\x01" + struct.pack("<HH", len(payload), len(payload) ^ 0xffff) + payload
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(data) + 25) + data +
            struct.pack("<II", zlib.crc32(payload), len(payload)))


def bam_stream(batches, schema: pa.Schema, tag_fields, zero_based: bool, overrides=None) -> bytes:
    """The uncompressed BAM byte stream write_bam_stream produces: header, then every record in arrival order."""
    text, names, lens = build_bam_header(schema, overrides)
    return bam_header_bytes(text, names, lens) + b"".join(encode_batch(b, names, tag_fields, zero_based) for b in batches)


def write_bam(path, batches, schema: pa.Schema, tag_fields, zero_based: bool = True, overrides=None, level: int = 6) -> int:
    """== BamWriteExec::execute for a .bam path; returns the row count (the reference's single-row `count` batch)."""
    batches = list(batches)
    stream = bam_stream(batches, schema, tag_fields, zero_based, overrides)
    with open(path, "wb") as f:
        for i in range(0, len(stream), BGZF_BLOCK):
            f.write(bgzf_member(stream[i:i + BGZF_BLOCK], level))
        f.write(BGZF_EOF)
    return sum(b.num_rows for b in batches)


def inflate_bgzf(data: bytes) -> tuple[bytes, list]:
    """Strict BGZF reader for the tests: every member's header, BSIZE, CRC-32 and ISIZE are checked.  -> (stream, [isize...])."""
    out, sizes, off = [], [], 0
    while off < len(data):
        if data[off:off + 4] != b"\x1f\x8b\x08\x04" or data[off + 12:off + 16] != b"BC\x02\x00":
            raise ValueError(f"bad BGZF member header at {off}")
        xlen = struct.unpack_from("<H", data, off + 10)[0]
        bsize = struct.unpack_from("<H", data, off + 16)[0] + 1
        raw = zlib.decompress(data[off + 12 + xlen: off + bsize - 8], -15)
        crc, isize = struct.unpack_from("<II", data, off + bsize - 8)
        if zlib.crc32(raw) != crc or len(raw) != isize:
            raise ValueError(f"CRC/ISIZE mismatch in the member at {off}")
        out.append(raw); sizes.append(isize); off += bsize
    return b"".join(out), sizes
