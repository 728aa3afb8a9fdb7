"""oracle/bam_oracle.py -- python face of the CPU restatement (oracle/bam_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product path (datafusion-bio-formats_b200/) never imports this module.

Restates, besides driving the C scan:
  * tag registry lookup / hint grammar      (reference: bio-format-core/src/tag_registry.rs:131-792)
  * determine_schema                        (bio-format-bam/src/table_provider.rs:42-140)
  * tag type inference by sampling          (table_provider.rs:145-202, 447-494)
  * SAM header -> bio.bam.* schema metadata (bio-format-core/src/metadata.rs:321-485)
"""
from __future__ import annotations

import ctypes as C
import json
import os
import re
import struct
import subprocess
import zlib
from pathlib import Path

import numpy as np
import pyarrow as pa

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
LIB_PATH = HERE / "_build" / "libbam_oracle.so"

K_Int32, K_UInt32, K_Float32, K_Utf8, K_Binary = 1, 2, 3, 4, 5
K_ListInt8, K_ListUInt8, K_ListInt16, K_ListUInt16, K_ListInt32, K_ListUInt32, K_ListFloat32 = range(10, 17)

_KIND_BY_NAME = {
    "Int32": K_Int32, "UInt32": K_UInt32, "Float32": K_Float32, "Utf8": K_Utf8,
    "ListInt8": K_ListInt8, "ListUInt8": K_ListUInt8, "ListInt16": K_ListInt16, "ListUInt16": K_ListUInt16,
    "ListInt32": K_ListInt32, "ListUInt32": K_ListUInt32, "ListFloat32": K_ListFloat32,
}
_LIST_INNER = {
    K_ListInt8: pa.int8(), K_ListUInt8: pa.uint8(), K_ListInt16: pa.int16(), K_ListUInt16: pa.uint16(),
    K_ListInt32: pa.int32(), K_ListUInt32: pa.uint32(), K_ListFloat32: pa.float32(),
}
_SUBTYPE_KIND = {"c": K_ListInt8, "C": K_ListUInt8, "s": K_ListInt16, "S": K_ListUInt16,
                 "i": K_ListInt32, "I": K_ListUInt32, "f": K_ListFloat32}
_KIND_SUBTYPE = {v: k for k, v in _SUBTYPE_KIND.items()}


def kind_to_arrow(kind: int) -> pa.DataType:
    if kind == K_Int32: return pa.int32()
    if kind == K_UInt32: return pa.uint32()
    if kind == K_Float32: return pa.float32()
    if kind == K_Utf8: return pa.utf8()
    if kind == K_Binary: return pa.binary()
    return pa.list_(pa.field("item", _LIST_INNER[kind], True))   # tag_registry.rs:17-19


def build() -> Path:
    """Compiles the C restatement (building the checker is not using it)."""
    src = HERE / "bam_oracle.c"
    LIB_PATH.parent.mkdir(exist_ok=True)
    if not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["gcc", "-O2", "-std=gnu11", "-shared", "-fPIC", "-fvisibility=hidden",
                               "-o", str(LIB_PATH), str(src), "-lz"])
    return LIB_PATH


# ---------------------------------------------------------------------------------------------
# registry / hints / schema

def load_registry() -> dict:
    """63 SAM-spec tags: tag -> (sam_type, kind, description).  Data table generated from the
    reference by tools/gen_tag_registry.py (tag_registry.rs:131-684)."""
    reg = {}
    inc = ROOT / "datafusion-bio-formats_b200" / "csrc" / "tag_registry_data.inc"
    pat = re.compile(r'TAGDEF\("(..)", \'(.)\', K_(\w+), "(.*)"\)')
    for line in inc.read_text().splitlines():
        m = pat.match(line)
        if m:
            reg[m.group(1)] = (m.group(2), _KIND_BY_NAME[m.group(3)], m.group(4))
    assert len(reg) == 63
    return reg


def parse_tag_type_hints(hints):
    """tag_registry.rs:698-752."""
    out = {}
    for hint in hints:
        parts = hint.split(":")
        if len(parts) == 2:
            tag, t = parts
            if len(t) != 1: raise ValueError(f"Invalid tag type hint '{hint}': TYPE must be a single character")
            if t == "B": raise ValueError(f"Invalid tag type hint '{hint}': array type 'B' requires a subtype")
            if t not in "AcCsSiIfZH": raise ValueError(f"Invalid tag type hint '{hint}': unsupported SAM type '{t}'")
            kind = {"A": K_Utf8, "c": K_Int32, "s": K_Int32, "i": K_Int32, "C": K_UInt32, "S": K_UInt32,
                    "I": K_UInt32, "f": K_Float32, "Z": K_Utf8, "H": K_Utf8}[t]
            out[tag] = (t, kind)
        elif len(parts) == 3 and parts[1] == "B":
            tag, _, st = parts
            if len(st) != 1: raise ValueError(f"Invalid tag type hint '{hint}': array subtype must be a single character")
            if st not in _SUBTYPE_KIND: raise ValueError(f"Invalid tag type hint '{hint}': unsupported array subtype '{st}'")
            out[tag] = ("B", _SUBTYPE_KIND[st])
        else:
            raise ValueError(f"Invalid tag type hint '{hint}': expected 'TAG:TYPE' or 'TAG:B:SUBTYPE' format")
    return out


def format_sam_tag_type(sam_type: str, kind: int) -> str:
    """tag_registry.rs:65-73."""
    if sam_type == "B" and kind in _KIND_SUBTYPE:
        return "B:" + _KIND_SUBTYPE[kind]
    return sam_type


def header_metadata(text: str) -> dict:
    """metadata.rs:321-485 over the SAM header text (noodles_sam::Header semantics)."""
    md = {}
    sq, rg, pg, co = [], [], [], []
    for line in text.split("\n"):
        line = line.rstrip("\r")
        if not line.startswith("@"):
            continue
        kind = line[:3]
        if kind == "@CO":
            co.append(line[4:] if len(line) > 3 else "")
            continue
        fields = {}
        order = []
        for tok in line.split("\t")[1:]:
            if len(tok) >= 3 and tok[2] == ":":
                fields[tok[:2]] = tok[3:]
                order.append(tok[:2])
        if kind == "@HD":
            if "VN" in fields: md["bio.bam.file_format_version"] = fields["VN"]
            if "SO" in fields: md["bio.bam.sort_order"] = fields["SO"]
            if "GO" in fields: md["bio.bam.group_order"] = fields["GO"]
            if "SS" in fields: md["bio.bam.subsort_order"] = fields["SS"]
        elif kind == "@SQ":
            d = {"name": fields.get("SN", ""), "length": int(fields.get("LN", "0"))}
            other = {k: fields[k] for k in order if k not in ("SN", "LN")}
            if other: d["other_fields"] = other
            sq.append(d)
        elif kind == "@RG":
            d = {"id": fields.get("ID", "")}
            for key, tag in (("sample", "SM"), ("platform", "PL"), ("library", "LB"), ("description", "DS")):
                if tag in fields: d[key] = fields[tag]
            other = {k: fields[k] for k in order if k not in ("ID", "SM", "PL", "LB", "DS")}
            if other: d["other_fields"] = other
            rg.append(d)
        elif kind == "@PG":
            d = {"id": fields.get("ID", "")}
            for key, tag in (("name", "PN"), ("version", "VN"), ("command_line", "CL")):
                if tag in fields: d[key] = fields[tag]
            other = {k: fields[k] for k in order if k not in ("ID", "PN", "VN", "CL")}
            if other: d["other_fields"] = other
            pg.append(d)
    if sq: md["bio.bam.reference_sequences"] = json.dumps(sq, separators=(",", ":"))
    if rg: md["bio.bam.read_groups"] = json.dumps(rg, separators=(",", ":"))
    if pg: md["bio.bam.program_info"] = json.dumps(pg, separators=(",", ":"))
    if co: md["bio.bam.comments"] = json.dumps(co, separators=(",", ":"))
    return md


# ---------------------------------------------------------------------------------------------
# ctypes plumbing

class _OrcTag(C.Structure):
    _fields_ = [("tag", C.c_char * 2), ("kind", C.c_int32)]


class _OrcFilter(C.Structure):
    _fields_ = [("op", C.c_int32), ("column", C.c_int32), ("n_num", C.c_int32), ("num", C.c_double * 16),
                ("n_str", C.c_int32), ("str", C.c_char_p * 16)]


class _OrcArgs(C.Structure):
    _fields_ = [("zero_based", C.c_int32), ("binary_cigar", C.c_int32), ("end_zero_span_mode", C.c_int32),
                ("n_tags", C.c_int32), ("tags", C.POINTER(_OrcTag)),
                ("n_proj", C.c_int32), ("proj", C.POINTER(C.c_int32)),
                ("start_voffset", C.c_uint64), ("end_coff", C.c_uint64), ("max_records", C.c_int64),
                ("batch_rows", C.c_int32),
                ("region_mode", C.c_int32), ("region_ref", C.c_int32),
                ("region_start", C.c_int64), ("region_end", C.c_int64), ("stop_voffset", C.c_uint64),
                ("n_filters", C.c_int32), ("filters", C.POINTER(_OrcFilter))]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB_PATH))
        L.orc_open.restype = C.c_void_p
        L.orc_open.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.orc_close.argtypes = [C.c_void_p]
        L.orc_n_ref.argtypes = [C.c_void_p]
        L.orc_ref_name.restype = C.c_char_p
        L.orc_ref_name.argtypes = [C.c_void_p, C.c_int]
        L.orc_ref_len.argtypes = [C.c_void_p, C.c_int]
        L.orc_header_text.restype = C.c_void_p
        L.orc_header_text.argtypes = [C.c_void_p]
        L.orc_header_text_len.argtypes = [C.c_void_p]
        L.orc_first_record_voffset.restype = C.c_uint64
        L.orc_first_record_voffset.argtypes = [C.c_void_p]
        L.orc_scan.restype = C.c_void_p
        L.orc_scan.argtypes = [C.c_void_p, C.POINTER(_OrcArgs)]
        L.orc_result_free.argtypes = [C.c_void_p]
        L.orc_result_error.restype = C.c_char_p
        L.orc_result_error.argtypes = [C.c_void_p]
        L.orc_result_rows.restype = C.c_int64
        L.orc_result_rows.argtypes = [C.c_void_p]
        L.orc_result_ncols.argtypes = [C.c_void_p]
        L.orc_result_stat.restype = C.c_int64
        L.orc_result_stat.argtypes = [C.c_void_p, C.c_int]
        L.orc_col_kind.argtypes = [C.c_void_p, C.c_int]
        L.orc_col_id.argtypes = [C.c_void_p, C.c_int]
        L.orc_col_nulls.restype = C.c_int64
        L.orc_col_nulls.argtypes = [C.c_void_p, C.c_int]
        L.orc_col_buf.restype = C.c_void_p
        L.orc_col_buf.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int64)]
        L.orc_index_records.restype = C.c_int64
        L.orc_index_records.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
        _lib = L
    return _lib


OPS = {"=": 0, "!=": 1, "<": 2, "<=": 3, ">": 4, ">=": 5, "between": 6, "not_between": 7, "in": 8, "not_in": 9}
COLUMN_IDS = {"chrom": 1, "start": 2, "end": 3, "flags": 4, "mapping_quality": 6}


class OracleBam:
    """CPU restatement of BamTableProvider for one local BAM (table_provider.rs:381-529)."""

    CORE_FIELDS = [("name", pa.utf8(), True), ("chrom", pa.utf8(), True), ("start", pa.uint32(), True),
                   ("end", pa.uint32(), True), ("flags", pa.uint32(), False), ("cigar", pa.utf8(), False),
                   ("mapping_quality", pa.uint32(), False), ("mate_chrom", pa.utf8(), True),
                   ("mate_start", pa.uint32(), True), ("sequence", pa.utf8(), False),
                   ("quality_scores", pa.utf8(), False), ("template_length", pa.int32(), False)]

    def __init__(self, path, zero_based=True, tag_fields=None, binary_cigar=False, infer_tag_types=True,
                 infer_tag_sample_size=100, tag_type_hints=None, end_zero_span_mode=0):
        self.path = str(path)
        self.zero_based = zero_based
        self.tag_fields = list(tag_fields) if tag_fields is not None else None
        self.binary_cigar = binary_cigar
        self.end_zero_span_mode = end_zero_span_mode
        err = C.create_string_buffer(256)
        self._h = lib().orc_open(self.path.encode(), err, 256)
        if not self._h:
            raise IOError(err.value.decode())
        L = lib()
        self.ref_names = [L.orc_ref_name(self._h, i).decode() for i in range(L.orc_n_ref(self._h))]
        self.ref_lens = [L.orc_ref_len(self._h, i) for i in range(len(self.ref_names))]
        n = L.orc_header_text_len(self._h)
        self.header_text = C.string_at(L.orc_header_text(self._h), n).decode("utf-8", "replace").rstrip("\0")
        self.first_record_voffset = L.orc_first_record_voffset(self._h)
        hints = parse_tag_type_hints(tag_type_hints) if tag_type_hints else None
        reg = load_registry()
        inferred = None
        if infer_tag_types and self.tag_fields:
            unknown = [t for t in self.tag_fields if t not in reg]
            if unknown:
                inferred = self._infer(unknown, infer_tag_sample_size)
        self.tag_specs = []   # (tag, sam_type, kind, description)
        for t in (self.tag_fields or []):     # table_provider.rs:73-126: inferred > hints > registry > Utf8
            if inferred and t in inferred:
                st, kind = inferred[t]
                desc = reg[t][2] if t in reg else f"Tag type discovered from file ({st})"
            elif hints and t in hints:
                st, kind = hints[t]
                desc = reg[t][2] if t in reg else f"Tag type from user hint ({st})"
            elif t in reg:
                st, kind, desc = reg[t]
            else:
                st, kind, desc = "Z", K_Utf8, "Unknown tag"
            self.tag_specs.append((t, st, kind, desc))
        self.schema = self._schema()

    def close(self):
        if self._h:
            lib().orc_close(self._h)
            self._h = None

    def __del__(self):
        try: self.close()
        except Exception: pass

    def _schema(self) -> pa.Schema:
        fields = []
        for name, ty, nullable in self.CORE_FIELDS:
            if name == "cigar" and self.binary_cigar:
                ty = pa.binary()
            fields.append(pa.field(name, ty, nullable))
        for tag, st, kind, desc in self.tag_specs:
            md = {"bio.bam.tag.tag": tag, "bio.bam.tag.type": format_sam_tag_type(st, kind),
                  "bio.bam.tag.description": desc}
            fields.append(pa.field(tag, kind_to_arrow(kind), True, metadata=md))
        md = header_metadata(self.header_text)
        md["bio.coordinate_system_zero_based"] = "true" if self.zero_based else "false"
        if self.binary_cigar:
            md["bio.bam.binary_cigar"] = "true"
        return pa.schema(fields, metadata=md)

    def _infer(self, tags, sample_size):
        """discover_tags_from_stream (table_provider.rs:145-202) over the first sample_size records."""
        found = {}
        for count, aux in enumerate(self._iter_aux(sample_size)):
            if count >= sample_size:
                break
            for t in tags:
                if t in found or len(t.encode()) != 2:
                    continue
                v = aux.get(t)
                if v is not None:
                    found[t] = v
        return found

    def _iter_aux(self, limit):
        """Pure-python aux walk of the first `limit` records (independent of the C code)."""
        data = open(self.path, "rb").read()
        off, stream = 0, bytearray()
        def more():
            nonlocal off
            if off >= len(data): return False
            bsize = struct.unpack_from("<H", data, off + 16)[0] + 1
            xlen = struct.unpack_from("<H", data, off + 10)[0]
            stream.extend(zlib.decompress(data[off + 12 + xlen: off + bsize - 8], -15))
            off += bsize
            return True
        while len(stream) < 12:
            if not more(): return
        l_text = struct.unpack_from("<i", stream, 4)[0]
        while len(stream) < 12 + l_text:
            if not more(): return
        p = 8 + l_text
        n_ref = struct.unpack_from("<i", stream, p)[0]; p += 4
        for _ in range(n_ref):
            while len(stream) < p + 4: more()
            l = struct.unpack_from("<i", stream, p)[0]
            while len(stream) < p + 8 + l:
                if not more(): return
            p += 8 + l
        n = 0
        while n < limit:
            while len(stream) < p + 4:
                if not more(): return
            bs = struct.unpack_from("<i", stream, p)[0]
            while len(stream) < p + 4 + bs:
                if not more(): return
            rec = bytes(stream[p + 4: p + 4 + bs]); p += 4 + bs; n += 1
            l_name, n_cig, l_seq = rec[8], struct.unpack_from("<H", rec, 12)[0], struct.unpack_from("<i", rec, 16)[0]
            a = 32 + l_name + 4 * n_cig + (l_seq + 1) // 2 + l_seq
            aux = {}
            while a + 3 <= len(rec):
                tag, ty = rec[a:a + 2].decode("latin-1"), chr(rec[a + 2]); a += 3
                if ty in "AcC": sz = 1
                elif ty in "sS": sz = 2
                elif ty in "iIf": sz = 4
                elif ty in "ZH": sz = rec.index(b"\0", a) - a + 1
                elif ty == "B":
                    st = chr(rec[a]); cnt = struct.unpack_from("<i", rec, a + 1)[0]
                    sz = 5 + cnt * {"c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4}[st]
                else: break
                if tag not in aux:   # tag_registry.rs:772-792
                    if ty == "A": aux[tag] = ("A", K_Utf8)
                    elif ty in "cCsSi": aux[tag] = ("i", K_Int32)
                    elif ty == "I": aux[tag] = ("I", K_UInt32)
                    elif ty == "f": aux[tag] = ("f", K_Float32)
                    elif ty == "Z": aux[tag] = ("Z", K_Utf8)
                    elif ty == "H": aux[tag] = ("H", K_Utf8)
                    else: aux[tag] = ("B", _SUBTYPE_KIND[chr(rec[a])])
                a += sz
            yield aux

    # -------------------------------------------------------------------------------------
    def projected_schema(self, projection):
        if projection is None:
            return self.schema
        return pa.schema([self.schema.field(i) for i in projection], metadata=self.schema.metadata)

    def scan(self, projection=None, start_voffset=0, end_coff=0, max_records=0, batch_rows=0,
             region=None, filters=None, stop_voffset=0, want_stats=False):
        """One partition -> pyarrow.RecordBatch (projection order).  `region` = (mode, ref_id, start, end)."""
        L = lib()
        a = _OrcArgs()
        a.zero_based = int(self.zero_based); a.binary_cigar = int(self.binary_cigar)
        a.end_zero_span_mode = self.end_zero_span_mode
        tags = (_OrcTag * max(1, len(self.tag_specs)))()
        for i, (t, _st, kind, _d) in enumerate(self.tag_specs):
            tb = t.encode()
            tags[i].tag = tb[:2] if len(tb) >= 2 else tb.ljust(2, b"\0")
            tags[i].kind = kind
        a.n_tags = len(self.tag_specs); a.tags = tags
        if projection is None:
            a.n_proj = -1; proj = None
        else:
            proj = (C.c_int32 * max(1, len(projection)))(*projection)
            a.n_proj = len(projection); a.proj = proj
        a.start_voffset = start_voffset; a.end_coff = end_coff; a.max_records = max_records
        a.batch_rows = batch_rows; a.stop_voffset = stop_voffset
        if region is not None:
            a.region_mode, a.region_ref, a.region_start, a.region_end = region
        keep = []
        if filters:
            fl = (_OrcFilter * len(filters))()
            for i, (col, op, vals) in enumerate(filters):
                fl[i].op = OPS[op]; fl[i].column = COLUMN_IDS[col]
                if col == "chrom":
                    fl[i].n_str = len(vals)
                    for k, v in enumerate(vals):
                        b = v.encode(); keep.append(b); fl[i].str[k] = b
                else:
                    fl[i].n_num = len(vals)
                    for k, v in enumerate(vals): fl[i].num[k] = float(v)
            a.n_filters = len(filters); a.filters = fl
        r = L.orc_scan(self._h, C.byref(a))
        try:
            err = L.orc_result_error(r)
            if err:
                raise RuntimeError(err.decode())
            n = L.orc_result_rows(r)
            arrays = []
            schema = self.projected_schema(projection)
            if batch_rows > 0:
                batch = None      # timing mode: batches were built and dropped inside the C loop
            else:
                for i in range(L.orc_result_ncols(r)):
                    arrays.append(self._col_to_arrow(r, i, n))
            if batch_rows > 0:
                pass
            elif arrays:
                batch = pa.RecordBatch.from_arrays(arrays, schema=schema)
            else:
                batch = pa.RecordBatch.from_struct_array(pa.array([{}] * n, type=pa.struct([]))) if n else \
                    pa.RecordBatch.from_pylist([], schema=schema)
            stats = {"inflated_bytes": L.orc_result_stat(r, 0), "blocks": L.orc_result_stat(r, 1),
                     "batches": L.orc_result_stat(r, 2), "records_seen": L.orc_result_stat(r, 3),
                     "next_voffset": L.orc_result_stat(r, 4), "rows": n}
        finally:
            L.orc_result_free(r)
        return (batch, stats) if want_stats else batch

    @staticmethod
    def _buf(r, i, which):
        n = C.c_int64()
        p = lib().orc_col_buf(r, i, which, C.byref(n))
        if not p or n.value == 0:
            return pa.py_buffer(b"")
        return pa.py_buffer(C.string_at(p, n.value))

    def _col_to_arrow(self, r, i, n):
        L = lib()
        kind = L.orc_col_kind(r, i)
        nulls = L.orc_col_nulls(r, i)
        validity = self._buf(r, i, 0) if nulls else None
        ty = kind_to_arrow(kind)
        if kind in (K_Int32, K_UInt32, K_Float32):
            return pa.Array.from_buffers(ty, n, [validity, self._buf(r, i, 2)], null_count=nulls)
        if kind in (K_Utf8, K_Binary):
            return pa.Array.from_buffers(ty, n, [validity, self._buf(r, i, 1), self._buf(r, i, 2)], null_count=nulls)
        inner = _LIST_INNER[kind]
        data = self._buf(r, i, 2)
        child = pa.Array.from_buffers(inner, data.size // (inner.bit_width // 8), [None, data], null_count=0)
        return pa.Array.from_buffers(ty, n, [validity, self._buf(r, i, 1)], null_count=nulls, children=[child])

    def time_scan(self, projection=None, start_voffset=0, end_coff=0, max_records=0, batch_rows=8192):
        """Reference-shaped timed loop: builders flushed and dropped every batch_rows rows."""
        _b, st = self.scan(projection, start_voffset, end_coff, max_records, batch_rows, want_stats=True)
        return st

    def index_records(self, stride=0, cap=1 << 22):
        """Virtual offsets of record starts (every `stride`-th record; 0 = first start of each block)."""
        L = lib()
        voff = np.zeros(cap, dtype=np.uint64); idx = np.zeros(cap, dtype=np.uint64)
        nrec = C.c_int64()
        k = L.orc_index_records(self._h, stride, voff.ctypes.data, idx.ctypes.data, cap, C.byref(nrec))
        return voff[:k].copy(), idx[:k].copy(), nrec.value


def schema_equal(a: pa.Schema, b: pa.Schema) -> bool:
    """Schema equality with bio.bam.* JSON metadata compared as parsed JSON."""
    if len(a) != len(b):
        return False
    for fa, fb in zip(a, b):
        if fa.name != fb.name or fa.type != fb.type or fa.nullable != fb.nullable:
            return False
        if (fa.metadata or {}) != (fb.metadata or {}):
            return False
    ma = {k.decode(): v.decode() for k, v in (a.metadata or {}).items()}
    mb = {k.decode(): v.decode() for k, v in (b.metadata or {}).items()}
    if ma.keys() != mb.keys():
        return False
    for k in ma:
        va, vb = ma[k], mb[k]
        if va == vb:
            continue
        try:
            if json.loads(va) != json.loads(vb):
                return False
        except ValueError:
            return False
    return True
