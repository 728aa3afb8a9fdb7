"""GPU parity of the BGZF-FASTQ scan against oracle/fastq_oracle.py: the reference's fixtures, projections,
multi-chunk scans with carried tail records, block-range partitions with speculated starts + seam check, edge cases."""
import pyarrow as pa
import pytest

from conftest import GOLDEN, make_fastq_text, write_bgzf
from oracle.fastq_oracle import OracleFastq

pytestmark = pytest.mark.gpu
FQ = GOLDEN / "fastq"


def _table(batch):
    return pa.Table.from_batches([batch])


def _assert_equal(got: pa.Table, want: pa.Table, what=""):
    assert got.schema.equals(want.schema), f"{what}: {got.schema} vs {want.schema}"
    assert got.num_rows == want.num_rows, f"{what}: {got.num_rows} rows, want {want.num_rows}"
    for name in want.schema.names:
        assert got[name].combine_chunks().equals(want[name].combine_chunks()), f"{what}: column {name} differs"


@pytest.mark.parametrize("name,rows", [("sample.fastq.bgz", 2000), ("example.fastq.bgz", 200)])
@pytest.mark.parametrize("flags", [0, 4], ids=["cta_per_member", "warp_per_member"])
def test_fixtures(name, rows, flags):
    """parallel_read_test.rs:65: 2000 rows; every column equals the oracle's."""
    import bamscan
    o = OracleFastq(FQ / name)
    p = bamscan.FastqTableProvider(str(FQ / name), debug_flags=flags)
    got = p.scan().collect()
    assert got.num_rows == rows
    _assert_equal(got, _table(o.scan()), name)
    p.close()


@pytest.mark.parametrize("proj", [[0], [2], [0, 2], [3], [1, 0], [3, 3, 1], []])
def test_projection(proj):
    """projection_pushdown_test.rs:41-58,76-100,119-135: single columns, pairs, order, the empty projection keeps the row count."""
    import bamscan
    o = OracleFastq(FQ / "sample.fastq.bgz")
    p = bamscan.FastqTableProvider(str(FQ / "sample.fastq.bgz"))
    got = p.scan(proj).collect()
    if not proj:                 # zero-column batches keep their row count (collect() returns them as a list)
        assert sum(b.num_rows for b in got) == 2000
        p.close()
        return
    want = _table(o.scan(proj))
    assert got.num_rows == 2000
    if proj:
        assert got.column_names == want.column_names
        for i in range(len(proj)):
            assert got.column(i).combine_chunks().equals(want.column(i).combine_chunks()), f"projection {proj} column {i}"
    p.close()


@pytest.mark.parametrize("mode", [None, "reference"])
@pytest.mark.parametrize("partitions", [1, 2, 4, 8])
def test_partitions_concatenate_to_sequential_scan(partitions, mode):
    """parallel_read_test.rs:230-231: the rows of 1 and of N partitions are the same; here also in the same order, and every
    partition's speculated first record is proven by the seam check.  mode None: block ranges balanced by compressed bytes;
    "reference": the reference's own cut, by block count from the companion .gzi (physical_exec.rs:140-175)."""
    import bamscan
    o = OracleFastq(FQ / "sample.fastq.bgz")
    p = bamscan.FastqTableProvider(str(FQ / "sample.fastq.bgz"))
    plan = p.scan(None, None, None, target_partitions=partitions, partition_mode=mode)
    n = plan.output_partition_count()
    parts, stats = [], []
    for i in range(n):
        parts += list(plan.execute(i))
        stats.append(plan.last_stats)
    got = pa.Table.from_batches(parts, schema=plan.schema())
    _assert_equal(got, _table(o.scan()), f"{partitions} partitions")
    if n > 1:
        bamscan.check_partition_seams(stats)
    p.close()


@pytest.mark.parametrize("crlf,final_newline", [(False, True), (False, False), (True, True)])
def test_synthetic_multi_chunk_and_partitions(tmp_path, crlf, final_newline):
    """60 k reads (13 MB of text, ~200 BGZF members) in chunks of 1 MiB: records straddle members and chunks (carried tails);
    then 1 / 3 / 7 block-range partitions.  Quality lines starting with '@' or '+', tab-separated and empty descriptions,
    CRLF line ends, a last line without a newline."""
    import bamscan
    txt = make_fastq_text(60000, 5, crlf=crlf, final_newline=final_newline)
    path = tmp_path / "syn.fastq.bgz"
    write_bgzf(path, txt)
    o = OracleFastq(path)
    want = _table(o.scan())
    assert want.num_rows == 60000
    p = bamscan.FastqTableProvider(str(path), chunk_inflated_bytes=1 << 20)
    _assert_equal(p.scan().collect(), want, "multi-chunk")
    for tp in (3, 7):
        plan = p.scan(None, None, None, target_partitions=tp)
        parts, stats = [], []
        for i in range(plan.output_partition_count()):
            parts += list(plan.execute(i))
            stats.append(plan.last_stats)
        _assert_equal(pa.Table.from_batches(parts, schema=plan.schema()), want, f"{tp} partitions")
        bamscan.check_partition_seams(stats)
    p.close()
    # default chunking (one chunk) and sliced batches
    p = bamscan.FastqTableProvider(str(path), batch_rows=8192)
    batches = list(p.scan().execute(0))
    assert max(b.num_rows for b in batches) <= 8192
    _assert_equal(pa.Table.from_batches(batches), want, "batch_rows")
    p.close()


def test_malformed_records_fail_loudly(tmp_path):
    import bamscan
    for i, txt in enumerate([b"@a\nAC\n-\nII\n", b"a\nAC\n+\nII\n", b"@a\nAC\n+\nII\n@b\nAC\n"]):
        path = tmp_path / f"bad{i}.fastq.bgz"
        write_bgzf(path, txt)
        p = bamscan.FastqTableProvider(str(path))
        with pytest.raises(bamscan.BamScanError):
            p.scan().collect()
        p.close()


def test_device_handoff(tmp_path):
    import bamscan
    o = OracleFastq(FQ / "example.fastq.bgz")
    p = bamscan.FastqTableProvider(str(FQ / "example.fastq.bgz"))
    plan = p.scan()
    got = pa.Table.from_batches([b.to_host() for b in plan.execute_device(0)])
    _assert_equal(got, _table(o.scan()), "device hand-off")
    p.close()
