import os
import struct
import subprocess
import sys
import zlib
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "datafusion-bio-formats_b200"
for p in (str(ROOT), str(PKG)):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def build_bamgen() -> Path:
    src = ROOT / "tools" / "bamgen.cpp"
    out = ROOT / "tools" / "_build" / "bamgen"
    out.parent.mkdir(exist_ok=True)
    if not out.exists() or out.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-o", str(out), str(src), "-lz"])
    return out


_GEN_CACHE = {}


def gen_bam(tmpdir: Path, mode="short", reads=20000, seed=1, bai=False, unmapped=0, level=6) -> Path:
    key = (mode, reads, seed, bai, unmapped, level)
    if key in _GEN_CACHE and _GEN_CACHE[key].exists():
        return _GEN_CACHE[key]
    exe = build_bamgen()
    out = Path(tmpdir) / f"syn_{mode}_{reads}_{seed}_{int(bai)}_{unmapped}_{level}.bam"
    cmd = [str(exe), "--mode", mode, "--reads", str(reads), "--seed", str(seed), "--out", str(out), "--level", str(level)]
    if bai:
        cmd.append("--bai")
    if unmapped:
        cmd += ["--unmapped", str(unmapped)]
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    _GEN_CACHE[key] = out
    return out


@pytest.fixture(scope="session")
def syn_dir(tmp_path_factory):
    return tmp_path_factory.mktemp("syn")


# ---- hand-made BAM with the edge cases the reference pins through write->read round trips (sam_read_test.rs) ----
def _bgzf(payload: bytes) -> bytes:
    out = b""
    for i in range(0, max(1, len(payload)), 0xff00):
        chunk = payload[i:i + 0xff00]
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        cd = c.compress(chunk) + c.flush()
        out += b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(cd) + 25) + cd + struct.pack("<II", zlib.crc32(chunk), len(chunk))
    return out + bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def _rec(ref, pos, name, mapq, flag, cigar, seq, qual, nref, npos, tlen, aux=b""):
    codes = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
    cg = b"".join(struct.pack("<I", (l << 4) | "MIDNSHP=X".index(o)) for l, o in cigar)
    sb = bytearray((len(seq) + 1) // 2)
    for i, ch in enumerate(seq):
        sb[i >> 1] |= codes[ch] << (4 if i % 2 == 0 else 0)
    body = struct.pack("<iiBBHHHiiii", ref, pos, len(name) + 1, mapq, 4680, len(cigar), flag, len(seq), nref, npos, tlen) + name.encode() + b"\0" + cg + bytes(sb) + bytes(qual) + aux
    return struct.pack("<i", len(body)) + body


def make_edge_bam(path, missing_qual=False):
    text = "@HD\tVN:1.6\tSO:unsorted\n@SQ\tSN:chrA\tLN:1000\tM5:abc\n@SQ\tSN:chrB\tLN:2000\n@RG\tID:g1\tSM:s\tPL:ILLUMINA\tXX:y\n@PG\tID:p\tPN:prog\tVN:1\tCL:cmd \"q\"\n@CO\thello\tworld\n"
    hdr = b"BAM\1" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", 2)
    for n, l in (("chrA", 1000), ("chrB", 2000)):
        hdr += struct.pack("<i", len(n) + 1) + n.encode() + b"\0" + struct.pack("<i", l)
    aux1 = b"NMC\x05" + b"MDZ10A5\0" + b"XSc\xfe" + b"XFf" + struct.pack("<f", 1.5) + b"XBBS" + struct.pack("<iHH", 2, 7, 65535) + b"XAAq" + b"XIi" + struct.pack("<i", -70000) + b"XUI" + struct.pack("<I", 4000000000)
    recs = [
        _rec(0, 99, "r1", 255, 99, [(5, "S"), (10, "M"), (2, "I"), (3, "D"), (4, "N"), (1, "="), (1, "X")], "ACGTNACGTNACGTNACGT", range(19), 1, 199, -160, aux1),
        _rec(-1, -1, "*", 0, 4, [], "", [], -1, -1, 0),                       # name "*", unmapped, l_seq 0 (sam_read_test.rs:902-907)
        _rec(1, 0, "r3", 60, 16, [(268435455, "M")], "A", [40], 0, 5, 160),   # max CIGAR length, odd l_seq
        _rec(0, 5, "placed_no_cigar", 3, 133, [], "AC", [1, 2], 0, 5, 0),     # CIGAR-less placed read: end NULL (unpinned default)
    ]
    if missing_qual:   # SAM spec 4.2.3: SEQ present, QUAL '*' is stored as 0xFF bytes; char::from(q + 33) wraps to ' ' in a release build
        recs.append(_rec(1, 7, "no_qual", 20, 0, [(40, "M")], "ACGT" * 10, [0xFF] * 40, -1, -1, 0))
        floats = [1e-7, 3.4028235e38, 0.1, -1.4e-45, 16777216.0, 0.0, -0.0, 123456.789, 5e-324, float("inf")]
        aux = b"".join(b"Y" + bytes([ord("a") + i]) + b"f" + struct.pack("<f", v) for i, v in enumerate(floats))
        recs.append(_rec(0, 9, "floats", 1, 0, [(4, "M")], "ACGT", [30] * 4, -1, -1, 0, aux))
    path.write_bytes(_bgzf(hdr + b"".join(recs)))




def write_bgzf(path: Path, data: bytes, block=0xff00, level=6):
    """`data` as a BGZF file (members of <= `block` inflated bytes + the EOF marker)."""
    with open(path, "wb") as f:
        for i in list(range(0, len(data), block)) + [None]:
            chunk = b"" if i is None else data[i:i + block]
            c = zlib.compressobj(level, zlib.DEFLATED, -15)
            body = c.compress(chunk) + c.flush()
            f.write(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(body) + 25))
            f.write(body + struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk)))


def make_fastq_text(n_reads: int, seed: int, read_len=101, crlf=False, final_newline=True) -> bytes:
    """Synthetic FASTQ: Illumina-like names, a mix of records with / without description (space- and tab-separated),
    quality lines that often START with '@' or '+' (the characters a naive record synchroniser trips over)."""
    import random
    rng = random.Random(seed)
    eol = "\r\n" if crlf else "\n"
    out = []
    for i in range(n_reads):
        n = read_len if rng.random() < 0.8 else rng.randint(1, 2 * read_len)
        seq = "".join(rng.choice("ACGTN") for _ in range(n))
        first = rng.choice("@+II?") if n else ""
        qual = (first + "".join(chr(33 + rng.randint(2, 40)) for _ in range(n - 1))) if n else ""
        kind = rng.random()
        name = f"SIM{seed}.{i}"
        if kind < 0.5:
            d = f"@{name} HSQ:{rng.randint(1, 9)}:{i}/1"
        elif kind < 0.6:
            d = f"@{name}\tcomment@with+signs {i}"
        elif kind < 0.65:
            d = f"@{name} "                      # delimiter but empty description -> NULL
        else:
            d = f"@{name}"
        out.append(d + eol + seq + eol + "+" + eol + qual + eol)
    text = "".join(out)
    if not final_newline and text.endswith(eol):
        text = text[:-len(eol)]
    return text.encode()


# ---- BAI -> CSI (CSIv1) with the same bins, or re-binned one level deeper; used by the CSI planner / scan tests ----
def bai_to_csi(bai_path, csi_path, depth=5, bgzf=True):
    """Rewrite a BAI as a CSI with min_shift 14 and `depth` levels (5 = BAI's own scheme; 6 .. 9 = every bin moved depth - 5 levels
    down, what `samtools index -c -m 14` gives contigs above 512 Mb).  A bin's loffset is the linear-index entry of its first 16 KiB
    window, as htslib computes it when it writes a CSI."""
    import struct
    d = Path(bai_path).read_bytes()
    assert d[:4] == b"BAI\x01" and 5 <= depth <= 9
    n_ref, = struct.unpack_from("<I", d, 4)
    p = 8
    first = lambda l: ((1 << (3 * l)) - 1) // 7
    out = [b"CSI\x01", struct.pack("<iii", 14, depth, 0), struct.pack("<I", n_ref)]
    for _ in range(n_ref):
        n_bin, = struct.unpack_from("<I", d, p); p += 4
        bins = []
        for _ in range(n_bin):
            b, n_chunk = struct.unpack_from("<II", d, p); p += 8
            bins.append((b, d[p:p + 16 * n_chunk], n_chunk)); p += 16 * n_chunk
        n_intv, = struct.unpack_from("<I", d, p); p += 4
        intv = struct.unpack_from(f"<{n_intv}Q", d, p); p += 8 * n_intv
        body = []
        for b, chunks, n_chunk in bins:
            if b == 37450:
                nb, loff = first(depth + 1) + 1, 0
            else:
                level = max(l for l in range(6) if first(l) <= b)
                k = b - first(level)
                nb = first(level + depth - 5) + k
                win = k << (3 * (5 - level))
                loff = intv[win] if win < n_intv else (intv[-1] if n_intv else 0)
            body.append(struct.pack("<IQi", nb, loff, n_chunk) + chunks)
        out.append(struct.pack("<I", len(body)) + b"".join(body))
    out.append(d[p:p + 8])                                              # n_no_coor, when the BAI has it
    raw = b"".join(out)
    Path(csi_path).write_bytes(_bgzf(raw) if bgzf else raw)
    return Path(csi_path)
