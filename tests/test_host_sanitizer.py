"""CPU: the product's host code (BGZF walk, BAM header, tag inference, BAI / CSI / GZI parsers, planner) linked without the CUDA
engine (tests/native/host_planner_driver.cpp) and run under ASan + UBSan on the fixtures and on corrupted copies of them.
Every input file is untrusted: a parser may refuse it, it may not read or write out of bounds."""
import random
import shutil
import subprocess
from pathlib import Path

import pytest

from conftest import GOLDEN, bai_to_csi

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "datafusion-bio-formats_b200" / "csrc"


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    exe = tmp_path_factory.mktemp("hpd") / "host_planner_driver"
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-o", str(exe),
           str(ROOT / "tests" / "native" / "host_planner_driver.cpp")] + [str(CSRC / f) for f in ("host_file.cpp", "host_schema.cpp", "host_index.cpp", "host_plan.cpp")] + ["-lz"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and "sanitize" in r.stderr:
        pytest.skip("no sanitizer runtime for g++ here: " + r.stderr[-200:])
    assert r.returncode == 0, r.stderr[-2000:]
    return exe


def _run(exe, *args):
    r = subprocess.run([str(exe), *map(str, args)], capture_output=True, text=True)
    assert r.returncode == 0 and "Sanitizer" not in r.stderr and "runtime error" not in r.stderr, (args, r.returncode, r.stderr[:3000])
    return r.stdout


@pytest.mark.parametrize("name", ["multi_chrom.bam", "multi_chrom_large.bam", "bam_with_tags.bam", "10x_pbmc_tags.bam", "nanopore_custom_tags.bam", "no_coor_only.bam"])
def test_fixtures_plan_cleanly(driver, name, tmp_path):
    assert "accepted" in _run(driver, GOLDEN / name, "")                 # with the discovered BAI
    bam = tmp_path / name
    shutil.copy(GOLDEN / name, bam)
    bai_to_csi(str(GOLDEN / name) + ".bai", str(bam) + ".csi", depth=7)
    assert "accepted" in _run(driver, bam, "")                           # with a discovered CSI
    assert "accepted" in _run(driver, GOLDEN / "fastq" / "sample.fastq.bgz", "-", "fastq")


def _mutate(rng, b):
    b = bytearray(b)
    mode = rng.randrange(4)
    if mode == 0 and len(b) > 1:
        b = b[:rng.randrange(1, len(b))]
    elif mode == 1:
        for _ in range(rng.randrange(1, 8)):
            b[rng.randrange(len(b))] = rng.randrange(256)
    elif mode == 2:
        for _ in range(rng.randrange(1, 8)):
            b[rng.randrange(min(len(b), 700))] = rng.randrange(256)
    else:
        pos = rng.randrange(0, max(1, min(len(b) - 4, 300)))
        b[pos:pos + 4] = rng.randrange(2 ** 32).to_bytes(4, "little")
    return bytes(b)


def test_corrupt_inputs_never_trip_the_sanitizers(driver, tmp_path):
    rng = random.Random(9)
    bams = [GOLDEN / n for n in ("multi_chrom.bam", "bam_with_tags.bam", "10x_pbmc_tags.bam", "nanopore_custom_tags.bam", "no_coor_only.bam")]
    fq = GOLDEN / "fastq" / "sample.fastq.bgz"
    for it in range(80):
        d = tmp_path / f"i{it}"
        d.mkdir()
        src, what = rng.choice(bams), it % 4
        bam = d / "f.bam"
        if what == 0:                                                    # corrupt BAM, no index
            bam.write_bytes(_mutate(rng, src.read_bytes()))
            _run(driver, bam, "-")
        elif what == 1:                                                  # good BAM, corrupt BAI
            shutil.copy(src, bam)
            (d / "f.bam.bai").write_bytes(_mutate(rng, Path(str(src) + ".bai").read_bytes()))
            _run(driver, bam, "")
        elif what == 2:                                                  # good BAM, corrupt CSI
            shutil.copy(src, bam)
            c = bai_to_csi(str(src) + ".bai", d / "t.csi", depth=rng.choice([5, 6, 9]), bgzf=it % 8 == 2).read_bytes()   # plain or BGZF-wrapped
            (d / "f.bam.csi").write_bytes(_mutate(rng, c))
            _run(driver, bam, "")
        else:                                                            # FASTQ: corrupt file or corrupt GZI
            f = d / "f.fq.bgz"
            if it % 8 == 3:
                f.write_bytes(_mutate(rng, fq.read_bytes()))
                shutil.copy(str(fq) + ".gzi", str(f) + ".gzi")
            else:
                shutil.copy(fq, f)
                Path(str(f) + ".gzi").write_bytes(_mutate(rng, Path(str(fq) + ".gzi").read_bytes()))
            _run(driver, f, "-", "fastq")
        shutil.rmtree(d)


def test_corrupt_inflated_contents_never_trip_the_sanitizers(driver, tmp_path):
    """Same, one layer down: the INFLATED stream is damaged (l_text / n_ref / reference entries, record and aux bytes, truncation)
    and re-packed into valid BGZF members (CRC-32 correct), so the damage reaches the header parser and the aux walk of the tag
    inference instead of stopping at the member checksum."""
    import struct
    import zlib
    from conftest import _bgzf

    def inflate_all(b):
        out, off = bytearray(), 0
        while off < len(b):
            bs = struct.unpack_from("<H", b, off + 16)[0] + 1
            xlen = struct.unpack_from("<H", b, off + 10)[0]
            out += zlib.decompress(b[off + 12 + xlen:off + bs - 8], -15)
            off += bs
        return bytes(out)

    rng = random.Random(13)
    srcs = [inflate_all((GOLDEN / n).read_bytes()) for n in ("multi_chrom.bam", "bam_with_tags.bam", "10x_pbmc_tags.bam", "nanopore_custom_tags.bam")]
    f = tmp_path / "a.bam"
    for it in range(45):
        u = bytearray(rng.choice(srcs))
        l_text = struct.unpack_from("<i", u, 4)[0]
        mode = it % 3
        if mode == 0:
            for _ in range(rng.randrange(1, 4)):
                u[rng.choice([4, 5, 6, 7, 8 + l_text, 9 + l_text, 12 + l_text, 13 + l_text, 8 + l_text + rng.randrange(64)])] = rng.randrange(256)
        elif mode == 1:
            base = 12 + l_text
            for _ in range(rng.randrange(1, 12)):
                u[rng.randrange(base, min(len(u), base + 60000))] = rng.randrange(256)
        else:
            u = u[:rng.randrange(8, len(u))]
        f.write_bytes(_bgzf(bytes(u)))
        _run(driver, f, "-")
