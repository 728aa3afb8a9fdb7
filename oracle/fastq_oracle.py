"""CPU restatement of the reference's FASTQ read path -- TEST INFRASTRUCTURE ONLY (nothing under datafusion-bio-formats_b200/
imports it; only tests/ may).

Follows datafusion/bio-format-fastq/src/physical_exec.rs:393-468 (batch_producer: `fastq::io::Reader::read_record`, four
StringBuilders, empty description -> NULL) and table_provider.rs:22-32 (determine_schema).  The record reader itself lives in
a dependency that is not vendored in /root/reference: noodles-fastq 0.23.0 (Cargo.lock, git fork biodatageeks/noodles @42a3c016).
Its published algorithm, restated: a record is exactly four lines -- '@' + definition, sequence, '+' line (content ignored),
quality scores; a line ends at '\\n' and one preceding '\\r' is dropped; the definition is split at its first space or tab
into name and description; a missing '@' / '+' prefix is an error; an unterminated last line is a line; a trailing partial
record is an error.

Pinned on the reference's own fixtures and test facts (tests/test_fastq_oracle.py): sample.fastq.bgz has 2000 rows
(tests/parallel_read_test.rs:65,125,230), every name / sequence / quality non-empty (:77,:95-96), partition counts sum to the
sequential count (:231).  The space-vs-tab split rule and '\\r' handling are NOT pinned by any reference fixture (all of them
use single spaces and '\\n'): "parity unpinned" for those two edges.
"""
import zlib

import pyarrow as pa

SCHEMA = pa.schema([pa.field("name", pa.string(), False), pa.field("description", pa.string(), True),
                    pa.field("sequence", pa.string(), False), pa.field("quality_scores", pa.string(), False)])


def gunzip_members(data: bytes) -> bytes:
    """Concatenated gzip / BGZF members -> one byte string (what MultiGzDecoder / bgzf::Reader hand to the FASTQ reader)."""
    out = []
    while data:
        d = zlib.decompressobj(31)
        out.append(d.decompress(data))
        if not d.eof:
            raise ValueError("truncated gzip member")
        data = d.unused_data
    return b"".join(out)


def parse_records(text: bytes):
    """-> list of (name, description or None, sequence, quality) byte strings."""
    lines = text.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()                       # the text ended with a newline
    lines = [ln[:-1] if ln.endswith(b"\r") else ln for ln in lines]
    if len(lines) % 4:
        raise ValueError("truncated FASTQ record at end of input")
    recs = []
    for i in range(0, len(lines), 4):
        d, s, p, q = lines[i:i + 4]
        if not d.startswith(b"@"):
            raise ValueError(f"record {i // 4}: missing '@' prefix")
        if not p.startswith(b"+"):
            raise ValueError(f"record {i // 4}: missing '+' line")
        d = d[1:]
        k = next((j for j, b in enumerate(d) if b in (0x20, 0x09)), len(d))
        name, desc = d[:k], d[k + 1:]
        recs.append((name, desc if desc else None, s, q))
    return recs


class OracleFastq:
    def __init__(self, path):
        self.path = str(path)
        self.schema = SCHEMA
        with open(self.path, "rb") as f:
            self._text = gunzip_members(f.read())
        self._recs = None

    def records(self):
        if self._recs is None:
            self._recs = parse_records(self._text)
        return self._recs

    def scan(self, projection=None) -> pa.RecordBatch:
        recs = self.records()
        cols = [pa.array([r[0].decode() for r in recs], pa.string()),
                pa.array([None if r[1] is None else r[1].decode() for r in recs], pa.string()),
                pa.array([r[2].decode() for r in recs], pa.string()),
                pa.array([r[3].decode() for r in recs], pa.string())]
        idx = list(range(4)) if projection is None else list(projection)
        if not idx:
            return pa.RecordBatch.from_struct_array(pa.array([{}] * len(recs), pa.struct([])))
        return pa.RecordBatch.from_arrays([cols[i] for i in idx], schema=pa.schema([SCHEMA.field(i) for i in idx]))
