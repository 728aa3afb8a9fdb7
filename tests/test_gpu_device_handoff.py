"""device-resident hand-off.  bamscan_execute_device / bamscan_next_device export every batch as an
ArrowDeviceArray (Arrow C Device Data Interface) whose buffers are device pointers; the test pulls the buffers back with
cudaMemcpy and compares the batch with the oracle, and checks that nothing was copied to the host by the engine."""
import pyarrow as pa
import pytest

from conftest import GOLDEN, gen_bam

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,tags", [("multi_chrom_large.bam", ["NM", "MD", "RG"]), ("nanopore_custom_tags.bam", ["pa", "ns", "NM", "MD"])])
def test_device_batches_equal_oracle(name, tags):
    import bamscan
    from oracle.bam_oracle import OracleBam
    path = GOLDEN / name
    want = pa.Table.from_batches([OracleBam(str(path), tag_fields=tags).scan()])
    p = bamscan.BamTableProvider(str(path), None, True, tags, False, True, 100, None, index_path="")
    plan = p.scan(None, [], None)
    got = []
    for db in plan.execute_device(0):
        assert db.device_type == bamscan.ARROW_DEVICE_CUDA and db.device_id == 0
        got.append(db.to_host())
        db.release()
    t = pa.Table.from_batches(got)
    assert plan.last_stats["d2h_bytes"] == 0                     # the engine copied no Arrow bytes to the host
    assert t.num_rows == want.num_rows
    for col in want.schema.names:
        assert t[col].combine_chunks().equals(want[col].combine_chunks()), col
    p.close()


def test_device_batches_multi_chunk_projection(syn_dir):
    import bamscan
    from oracle.bam_oracle import OracleBam
    path = gen_bam(syn_dir, "short", 60000, seed=11)
    proj = [2, 0, 9, 12]
    want = pa.Table.from_batches([OracleBam(str(path), tag_fields=["NM"]).scan(projection=proj)])
    p = bamscan.BamTableProvider(str(path), None, True, ["NM"], False, True, 100, None, index_path="", chunk_inflated_bytes=4 << 20)
    plan = p.scan(proj, [], None)
    got = [db.to_host() for db in plan.execute_device(0)]
    assert len(got) > 1
    t = pa.Table.from_batches(got)
    assert t.schema.names == want.schema.names
    for col in want.schema.names:
        assert t[col].combine_chunks().equals(want[col].combine_chunks()), col
    p.close()
