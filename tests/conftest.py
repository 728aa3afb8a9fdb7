import os
import subprocess
import sys
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
