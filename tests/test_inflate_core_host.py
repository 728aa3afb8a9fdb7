"""CPU: the per-lane code of the CTA-per-member inflate kernel (csrc/inflate_cta_core.h: speculative sub-stream decode with
warm-up, restart rounds, emit, LZ77 resolve) compiled for the host by tools/inflate_sim.cpp -- the lanes of a CTA run in a loop --
and compared with zlib on every BGZF member of the reference's fixtures and of synthetic files at three DEFLATE levels.
The GPU parity tests check the kernel itself; this pins the shared algorithm where no GPU exists."""
import re
import subprocess
from pathlib import Path

import pytest

from conftest import GOLDEN, gen_bam

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def inflate_sim(tmp_path_factory):
    exe = tmp_path_factory.mktemp("isim") / "inflate_sim"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", str(exe), str(ROOT / "tools" / "inflate_sim.cpp"), "-lz"])
    return exe


def _run(exe, path, *args):
    r = subprocess.run([str(exe), str(path), *args], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(r"members=(\d+) bad=(\d+) refused\(sub-table\)=(\d+)", r.stdout)
    assert m, r.stdout
    return tuple(int(x) for x in m.groups())


@pytest.mark.parametrize("name", ["multi_chrom.bam", "multi_chrom_large.bam", "bam_with_tags.bam", "10x_pbmc_tags.bam",
                                  "nanopore_custom_tags.bam", "no_coor_only.bam", "fastq/sample.fastq.bgz", "fastq/example.fastq.bgz"])
@pytest.mark.parametrize("lanes", [256, 64])
def test_host_model_equals_zlib_on_fixtures(inflate_sim, name, lanes):
    members, bad, _refused = _run(inflate_sim, GOLDEN / name, "--lanes", str(lanes))
    assert members > 0 and bad == 0


@pytest.mark.parametrize("mode,reads,level", [("short", 30000, 1), ("short", 30000, 6), ("short", 30000, 9), ("long", 300, 6)])
def test_host_model_equals_zlib_on_synthetic_files(inflate_sim, syn_dir, mode, reads, level):
    path = gen_bam(syn_dir, mode, reads, seed=3, level=level)
    members, bad, refused = _run(inflate_sim, path)
    assert members > 20 and bad == 0 and refused <= members // 10     # refused members go to the warp-per-member kernel on the GPU
    # a second sub-stream geometry (shorter spans, shorter warm-up) must decode the same bytes
    members2, bad2, _ = _run(inflate_sim, path, "--span-bytes", "2048", "--overlap", "256", "--max-members", "200")
    assert members2 > 0 and bad2 == 0


def test_corrupt_payloads_under_address_sanitizer(tmp_path):
    """compute-sanitizer is closed on the GPU pool; the shared per-lane code runs here under ASan + UBSan instead: byte flips inside
    the DEFLATE payloads (half of them in the block-header bytes that describe the Huffman tables) must end as a clean refusal
    or as zlib's own output -- no out-of-bounds access, no undefined shift, no disagreement with zlib."""
    import random
    import struct
    exe = tmp_path / "inflate_sim_asan"
    r = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-o", str(exe),
                        str(ROOT / "tools" / "inflate_sim.cpp"), "-lz"], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("no sanitizer runtime for g++ here: " + r.stderr[-200:])
    src = (GOLDEN / "multi_chrom.bam").read_bytes()
    members, off = [], 0
    while off < len(src):
        bs = struct.unpack_from("<H", src, off + 16)[0] + 1
        xlen = struct.unpack_from("<H", src, off + 10)[0]
        if off + bs - 8 > off + 12 + xlen + 2:
            members.append((off + 12 + xlen, off + bs - 8))
        off += bs
    rng = random.Random(5)
    f = tmp_path / "f.bin"
    for it in range(40):
        b = bytearray(src)
        for _ in range(rng.randrange(1, 6)):
            lo, hi = rng.choice(members)
            pos = lo + (rng.randrange(min(hi - lo, 80)) if rng.random() < 0.5 else rng.randrange(hi - lo))
            b[pos] = rng.randrange(256)
        f.write_bytes(bytes(b))
        r = subprocess.run([str(exe), str(f)], capture_output=True, text=True)
        assert r.returncode == 0 and "AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr, (it, r.returncode, r.stderr[:2000])
