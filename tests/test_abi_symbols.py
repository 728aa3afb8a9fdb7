"""CPU: the C-ABI library loads and exports every symbol include/bamscan.h declares; planning calls work without a GPU;
the scan itself refuses to run without one (no CPU fallback)."""
import ctypes
import re

import pytest

from conftest import GOLDEN, ROOT


def test_every_declared_symbol_is_exported():
    import bamscan
    header = (ROOT / "include" / "bamscan.h").read_text()
    declared = sorted(set(re.findall(r"\b(bamscan_[a-z_]+)\s*\(", header)))
    assert len(declared) >= 16
    lib = bamscan.load_library()
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/bamscan.h but not exported by libbamscan.so"
    assert set(bamscan.EXPORTED_SYMBOLS) == set(declared)
    assert b"sm_100a" in lib.bamscan_version()


def test_header_cites_reference_interfaces():
    header = (ROOT / "include" / "bamscan.h").read_text()
    for cite in ("table_provider.rs:381-390", "table_provider.rs:941-962", "table_provider.rs:964-1115", "physical_exec.rs:108-172"):
        assert cite in header


def test_library_links_no_torch_and_oracle():
    import subprocess, bamscan
    out = subprocess.check_output(["ldd", str(bamscan.LIB_PATH)]).decode()
    assert "torch" not in out and "oracle" not in out


def test_scan_refuses_without_gpu():
    import bamscan
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    p = bamscan.BamTableProvider(str(GOLDEN / "multi_chrom.bam"))
    plan = p.scan(None, [], None)
    with pytest.raises(bamscan.BamScanError) as e:
        list(plan.execute(0))
    assert e.value.code == -4 and "no CPU path" in str(e.value)


def test_writer_validates_the_schema_and_refuses_without_gpu(tmp_path):
    """The write path checks the input schema on the host like the reference's column look-ups
    (sam_record_serializer.rs:27-37, 248-279) and, like the scan, has no CPU path behind it."""
    import pyarrow as pa
    import bamscan
    from oracle.bam_oracle import OracleBam
    o = OracleBam(str(GOLDEN / "multi_chrom.bam"), tag_fields=["NM"])
    out = tmp_path / "o.bam"
    no_flags = pa.schema([f for f in o.schema if f.name != "flags"], metadata=o.schema.metadata)
    with pytest.raises(bamscan.BamScanError, match="Required column 'flags' not found") as e:
        bamscan.BamWriteExec(str(out), no_flags, ["NM"]).execute([])
    assert e.value.code == -7
    wrong = pa.schema([pa.field("start", pa.int64()) if f.name == "start" else f for f in o.schema], metadata=o.schema.metadata)
    with pytest.raises(bamscan.BamScanError, match="Column 'start' must be UInt32"):
        bamscan.BamWriteExec(str(out), wrong, ["NM"]).execute([])
    bad_tag = pa.schema(list(o.schema)[:12] + [pa.field("NM", pa.int32(), True, {"bio.bam.tag.tag": "NM", "bio.bam.tag.type": "Z"})])
    with pytest.raises(bamscan.BamScanError, match="Tag value type mismatch"):
        bamscan.BamWriteExec(str(out), bad_tag, ["NM"]).execute([])
    with pytest.raises(bamscan.BamScanError, match="plain SAM output is out of scope"):
        bamscan.BamWriteExec(str(tmp_path / "o.sam"), o.schema, ["NM"])
    try:
        import torch
        if torch.cuda.is_available():
            return
    except ImportError:
        pass
    with pytest.raises(bamscan.BamScanError, match="no CPU path") as e:
        bamscan.BamWriteExec(str(out), o.schema, ["NM"]).execute([])
    assert e.value.code == -4 and not out.exists()


def test_cpp_mirror_header_compiles_and_plans(tmp_path):
    """include/bamscan.hpp (the C++ face of BamTableProvider / BamExec over the C ABI: the reference's host language, Rust, is not
    available here) compiles with -Wall against the library and reproduces the planning pins: schema, pushdown classes, EmptyExec
    for an unsatisfiable conjunction, partition counts, describe rows, error codes."""
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    lib = root / "datafusion-bio-formats_b200"
    exe = tmp_path / "cpp_mirror_check"
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", f"-I{root / 'include'}", "-o", str(exe), str(root / "tests" / "native" / "cpp_mirror_check.cpp"),
                        str(lib / "libbamscan.so"), f"-Wl,-rpath,{lib}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([str(exe), str(root / "tests" / "golden" / "multi_chrom.bam"), str(root / "tests" / "golden" / "10x_pbmc_tags.bam"),
                        str(root / "tests" / "golden" / "fastq" / "sample.fastq.bgz")], capture_output=True, text=True)
    assert r.returncode == 0 and "cpp mirror ok" in r.stdout, r.stdout + r.stderr


def test_c_abi_header_is_plain_c99(tmp_path):
    """The drop-in boundary is a C ABI: include/bamscan.h must parse as strict C99 (what cgo / bindgen / ctypes generators read)."""
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    src = tmp_path / "c_abi.c"
    src.write_text('#include "bamscan.h"\nint main(void) { return bamscan_version() == 0; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", f"-I{root / 'include'}", "-fsyntax-only", str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
