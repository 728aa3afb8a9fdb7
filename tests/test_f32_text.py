"""csrc/f32_text.h (what the GPU runs for a float tag in a Utf8 column) against the oracle's rule, on the CPU:
tests/native/f32_text_check.cpp compiles both and compares ~300 k bit patterns incl. subnormals, powers of two and infinities."""
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_f32_text_matches_oracle_rule(tmp_path):
    exe = tmp_path / "f32chk"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", str(exe), str(ROOT / "tests" / "native" / "f32_text_check.cpp")])
    out = subprocess.check_output([str(exe), "120000"]).decode()
    assert "bad=0" in out, out
