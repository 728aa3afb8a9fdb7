"""CPU: the FASTQ oracle against the reference's own fixtures and test facts; the FASTQ entry of the C ABI plans without a GPU."""
import pyarrow as pa
import pytest

from conftest import GOLDEN, make_fastq_text, write_bgzf
from oracle.fastq_oracle import OracleFastq, parse_records

FQ = GOLDEN / "fastq"


def test_sample_fixture_facts():
    """bio-format-fastq/tests/parallel_read_test.rs:65,77,95-96: 2000 rows, no empty name / sequence / quality."""
    o = OracleFastq(FQ / "sample.fastq.bgz")
    b = o.scan()
    assert b.num_rows == 2000 and b.schema.names == ["name", "description", "sequence", "quality_scores"]
    assert b.schema.field("description").nullable and not b.schema.field("name").nullable
    assert all(len(x) > 0 for x in b.column(0).to_pylist()) and all(len(x) > 0 for x in b.column(2).to_pylist())
    assert b.column(0)[0].as_py() == "ERR194146.812444541" and b.column(1)[0].as_py() == "HSQ1008:141:D0CC8ACXX:2:1204:13288:78171/2"
    assert [len(s) for s in b.column(2).to_pylist()] == [len(q) for q in b.column(3).to_pylist()]
    assert OracleFastq(FQ / "example.fastq.bgz").scan().num_rows == 200


def test_record_rules():
    """projection_pushdown_test.rs:7-20,55-57,98-100,133-135 (inline sample) + the edges the restatement defines."""
    txt = b"@read1\nATCGATCGATCG\n+\nIIIIIIIIIIII\n@read2 d e\nGCTAGCTAGCTA\n+read2\nJJJJJJJJJJJJ\n@read3\tx\r\nAC\r\n+\r\n@K"
    r = parse_records(txt)
    assert r == [(b"read1", None, b"ATCGATCGATCG", b"IIIIIIIIIIII"), (b"read2", b"d e", b"GCTAGCTAGCTA", b"JJJJJJJJJJJJ"), (b"read3", b"x", b"AC", b"@K")]
    with pytest.raises(ValueError):
        parse_records(b"read1\nA\n+\nI\n")
    with pytest.raises(ValueError):
        parse_records(b"@r\nA\n-\nI\n")
    with pytest.raises(ValueError):
        parse_records(b"@r\nA\n+\nI\n@s\nA\n")


def test_abi_plans_fastq_without_gpu(tmp_path):
    import bamscan
    p = bamscan.FastqTableProvider(str(FQ / "sample.fastq.bgz"))
    assert p.schema().equals(OracleFastq(FQ / "sample.fastq.bgz").schema)
    assert p.supports_filters_pushdown([("name", "=", ["x"])]) == ["Unsupported"]
    plan = p.scan([2], None, None, target_partitions=4)
    assert plan.output_partition_count() == 4 and plan.schema().names == ["sequence"]
    ranges = [plan.partition_ranges(i) for i in range(4)]
    assert ranges[0][0]["exact_start"] and not ranges[1][0]["exact_start"]
    assert [r[0]["block_begin"] for r in ranges[1:]] == [r[0]["block_end"] for r in ranges[:-1]]
    p.close()
    # not BGZF: refused
    bad = tmp_path / "plain.fastq"
    bad.write_bytes(b"@r\nA\n+\nI\n")
    with pytest.raises(bamscan.BamScanError):
        bamscan.FastqTableProvider(str(bad))


def test_generator_round_trips(tmp_path):
    txt = make_fastq_text(3000, 7)
    path = tmp_path / "g.fastq.bgz"
    write_bgzf(path, txt)
    o = OracleFastq(path)
    assert o.scan().num_rows == 3000 and o._text == txt
    assert o.scan().column(1).null_count > 0
