#!/usr/bin/env python3
"""Generates tests/golden/golden_facts.json: per-fixture facts computed by an INDEPENDENT pure-python restatement
(stdlib zlib + struct only, no code shared with oracle/ or the product).

The fixtures are the reference's own test inputs (datafusion/bio-format-bam/tests/*.bam, copied byte-for-byte; data,
not source).  The facts restate what the reference's tests pin (row counts 421 = 160+159+102, 4277 = 1662+1694+921,
20 / 14 / 10 / 2 rows: indexed_read_test.rs:76-77,108,121; indexed_read_large_test.rs:63,85,95; tag_tests.rs) plus
full-column aggregates (sums, byte counts, sha256 of concatenated columns) following the column rules of
physical_exec.rs:412-528 and alignment_utils.rs:667-701, with 0-based `start`.  They agree with SURVEY.md App. C.

    python tests/golden/make_golden.py        # rewrites golden_facts.json (deterministic)
"""
import hashlib
import json
import struct
import zlib
from pathlib import Path

HERE = Path(__file__).resolve().parent
FIXTURES = ["multi_chrom.bam", "multi_chrom_large.bam", "nanopore_custom_tags.bam", "bam_with_tags.bam",
            "10x_pbmc_tags.bam", "no_coor_only.bam"]
SEQ = "=ACMGRSVTWYHKDBN"
OPS = "MIDNSHP=X"


def inflate_all(data: bytes):
    off, out, isizes = 0, bytearray(), []
    while off < len(data):
        assert data[off:off + 4] == b"\x1f\x8b\x08\x04"
        xlen = struct.unpack_from("<H", data, off + 10)[0]
        x, bsize = 0, None
        while x < xlen:
            si1, si2, slen = struct.unpack_from("<BBH", data, off + 12 + x)
            if (si1, si2, slen) == (66, 67, 2):
                bsize = struct.unpack_from("<H", data, off + 12 + x + 4)[0]
            x += 4 + slen
        total = bsize + 1
        raw = zlib.decompress(data[off + 12 + xlen: off + total - 8], -15)
        crc, isize = struct.unpack_from("<II", data, off + total - 8)
        assert zlib.crc32(raw) == crc and len(raw) == isize
        out += raw
        isizes.append(isize)
        off += total
    return bytes(out), isizes


def parse(path: Path):
    s, isizes = inflate_all(path.read_bytes())
    assert s[:4] == b"BAM\x01"
    l_text = struct.unpack_from("<i", s, 4)[0]
    p = 8 + l_text
    n_ref = struct.unpack_from("<i", s, p)[0]; p += 4
    refs = []
    for _ in range(n_ref):
        l = struct.unpack_from("<i", s, p)[0]
        refs.append(s[p + 4: p + 4 + l - 1].decode()); p += 8 + l
    recs = []
    while p < len(s):
        bs = struct.unpack_from("<i", s, p)[0]
        r = s[p + 4: p + 4 + bs]; p += 4 + bs
        ref, pos, l_name, mapq, _bin, n_cig, flag, l_seq, nref, npos, tlen = struct.unpack_from("<iiBBHHHiiii", r, 0)
        o = 32
        name = r[o:o + l_name - 1].decode(); o += l_name
        cig = struct.unpack_from(f"<{n_cig}I", r, o); o += 4 * n_cig
        seqb = r[o:o + (l_seq + 1) // 2]; o += (l_seq + 1) // 2
        qual = r[o:o + l_seq]; o += l_seq
        span = sum(c >> 4 for c in cig if (c & 15) in (0, 2, 3, 7, 8))
        seq = "".join(SEQ[(seqb[i >> 1] >> 4) if i % 2 == 0 else (seqb[i >> 1] & 15)] for i in range(l_seq))
        recs.append(dict(
            name=name, chrom=refs[ref] if ref >= 0 else None, start=pos if pos >= 0 else None,
            end=(pos + span) if (pos >= 0 and span > 0) else None, flags=flag,
            cigar="".join(f"{c >> 4}{OPS[c & 15]}" for c in cig), mapq=mapq,
            mate_chrom=refs[nref] if nref >= 0 else None, mate_start=npos if npos >= 0 else None,
            seq=seq, qual="".join(chr(q + 33) for q in qual), tlen=tlen, n_aux_bytes=bs - o))
    return refs, recs, isizes


def sha(s: str) -> str:
    return hashlib.sha256(s.encode()).hexdigest()[:16]


def facts(path: Path) -> dict:
    refs, recs, isizes = parse(path)
    per_chrom = {}
    for r in recs:
        per_chrom[str(r["chrom"])] = per_chrom.get(str(r["chrom"]), 0) + 1
    pick = lambda r: {k: r[k] for k in ("name", "chrom", "start", "end", "flags", "cigar", "mapq", "mate_chrom", "mate_start", "tlen")}
    return {
        "sha256_16": hashlib.sha256(path.read_bytes()).hexdigest()[:16],
        "size": path.stat().st_size, "bgzf_isizes": isizes, "n_refs": len(refs), "records": len(recs), "per_chrom": per_chrom,
        "sum_start": sum(r["start"] for r in recs if r["start"] is not None),
        "sum_end": sum(r["end"] for r in recs if r["end"] is not None),
        "null_end": sum(1 for r in recs if r["end"] is None), "null_start": sum(1 for r in recs if r["start"] is None),
        "sum_flags": sum(r["flags"] for r in recs), "sum_mapq": sum(r["mapq"] for r in recs),
        "sum_mate_start": sum(r["mate_start"] for r in recs if r["mate_start"] is not None),
        "sum_tlen": sum(r["tlen"] for r in recs),
        "bytes_name": sum(len(r["name"]) for r in recs), "bytes_cigar": sum(len(r["cigar"]) for r in recs),
        "bytes_seq": sum(len(r["seq"]) for r in recs),
        "sha_seq": sha("".join(r["seq"] for r in recs)), "sha_qual": sha("".join(r["qual"] for r in recs)),
        "sha_cigar": sha(",".join(r["cigar"] for r in recs)), "sha_name": sha(",".join(r["name"] for r in recs)),
        "first": pick(recs[0]), "last": pick(recs[-1]),
        "first_seq_prefix": recs[0]["seq"][:20], "first_qual_prefix": recs[0]["qual"][:20],
    }


if __name__ == "__main__":
    out = {f: facts(HERE / f) for f in FIXTURES}
    (HERE / "golden_facts.json").write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")
    for f, v in out.items():
        print(f, v["records"], v["per_chrom"], v["sum_start"], v["sha_seq"])
