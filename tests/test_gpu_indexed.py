"""GPU parity of the indexed (BAI) scan: BASELINE config 4 and the reference's indexed_read_test.rs /
indexed_read_large_test.rs pins.  Expected rows come from the oracle applying the reference's row rules
(physical_exec.rs:1036-1356) over the very chunk ranges the plan exposes."""
import collections

import pyarrow as pa
import pyarrow.compute as pc
import pytest

from conftest import GOLDEN, gen_bam

pytestmark = pytest.mark.gpu

ACCESSOR_COLS = {"chrom", "start", "end", "mapping_quality", "flags"}


def _provider(path, **kw):
    import bamscan
    return bamscan.BamTableProvider(str(path), None, kw.pop("zero_based", True), kw.pop("tag_fields", None), False, True, 100, None, **kw)


def _oracle(path, **kw):
    from oracle.bam_oracle import OracleBam
    return OracleBam(str(path), **kw)


def expected_partition(o, plan, p, filters, projection=None):
    import bamscan
    pushable = [f for f, how in zip(filters, plan._provider.supports_filters_pushdown(filters)) if how == "Inexact" and f[1] != "other"] if filters else []
    residual = [f for f in pushable if f[0] in ACCESSOR_COLS]          # other columns always pass at record level
    # keep only filters the record-level evaluator can push (can_push_down_record_filter)
    out = []
    for r in plan.partition_ranges(p):
        b = o.scan(projection=projection, start_voffset=r["start_voffset"], stop_voffset=r["stop_voffset"],
                   region=(r["region_mode"], r["region_ref"], r["region_start"], r["region_end"]), filters=residual)
        out.append(b)
    return out


def run_partition(plan, p):
    return list(plan.execute(p))


def assert_same(got_batches, want_batches, ctx):
    if not want_batches:
        assert sum(b.num_rows for b in got_batches) == 0, ctx
        return
    want = pa.Table.from_batches(want_batches)
    if want.num_rows == 0:
        assert sum(b.num_rows for b in got_batches) == 0, ctx
        return
    got = pa.Table.from_batches(got_batches, schema=want.schema) if got_batches else want.slice(0, 0)
    assert got.num_rows == want.num_rows, f"{ctx}: {got.num_rows} != {want.num_rows}"
    for name in want.schema.names:
        assert got[name].combine_chunks().equals(want[name].combine_chunks()), f"{ctx}: column {name}"


@pytest.mark.parametrize("tp", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("name,total,per", [("multi_chrom.bam", 421, (160, 159, 102)), ("multi_chrom_large.bam", 4277, (1662, 1694, 921))])
def test_indexed_full_scan_counts_are_partition_invariant(name, total, per, tp):
    # indexed_read_test.rs:76-77,108,121,298-319 ; indexed_read_large_test.rs:63,85,95
    path = GOLDEN / name
    p = _provider(path, tag_fields=["NM"])
    o = _oracle(path, tag_fields=["NM"])
    plan = p.scan(None, [], None, target_partitions=tp)
    n = plan.output_partition_count()
    assert 1 <= n <= tp
    rows, chroms = 0, collections.Counter()
    for i in range(n):
        got = run_partition(plan, i)
        assert_same(got, expected_partition(o, plan, i, []), f"{name} tp={tp} partition {i}")
        for b in got:
            rows += b.num_rows
            chroms.update(b.column("chrom").to_pylist())
    assert rows == total and (chroms["chr1"], chroms["chr2"], chroms["chrX"]) == per


@pytest.mark.parametrize("filters,want_rows", [
    ([("chrom", "=", ["chr1"])], 160),
    ([("chrom", "in", ["chr1", "chrX"])], 262),
    ([("chrom", "=", ["chr2"]), ("mapping_quality", ">=", [30])], None),
    ([("chrom", "=", ["chr1"]), ("start", "between", [55000000, 55001000])], None),
    ([("chrom", "=", ["chrX"]), ("start", ">=", [48001000]), ("end", "<=", [48001500]), ("flags", "not_in", [99, 147])], None),
    ([("chrom", "=", ["chr1"]), ("start", ">", [55000500]), ("name", "=", ["zzz"]), ("chrom", "!=", ["chr2"])], None),
])
@pytest.mark.parametrize("tp", [1, 4])
def test_region_queries_match_manual_filter_of_full_scan(filters, want_rows, tp):
    # indexed_read_test.rs:208-237: indexed result == the same predicate applied to a full scan
    path = GOLDEN / "multi_chrom.bam"
    p = _provider(path)
    o = _oracle(path)
    plan = p.scan(None, filters, None, target_partitions=tp)
    got = []
    for i in range(plan.output_partition_count()):
        g = run_partition(plan, i)
        assert_same(g, expected_partition(o, plan, i, filters), f"{filters} tp={tp} p{i}")
        got += g
    full = pa.Table.from_batches([o.scan()])
    mask = None
    for col, op, vals in filters:          # manual filter, SQL semantics, on the pushed-down (accessor) columns only
        if col not in ACCESSOR_COLS:
            continue
        c = full[col]
        if op == "=": m = pc.equal(c, vals[0])
        elif op == "!=": m = pc.not_equal(c, vals[0])
        elif op == ">=": m = pc.greater_equal(c, vals[0])
        elif op == ">": m = pc.greater(c, vals[0])
        elif op == "<=": m = pc.less_equal(c, vals[0])
        elif op == "in": m = pc.is_in(c, value_set=pa.array(vals))
        elif op == "not_in": m = pc.invert(pc.is_in(c, value_set=pa.array(vals, type=c.type)))
        elif op == "between": m = pc.and_(pc.greater_equal(c, vals[0]), pc.less_equal(c, vals[1]))
        mask = m if mask is None else pc.and_(mask, m)
    manual = full.filter(mask)
    n_got = sum(b.num_rows for b in got)
    if want_rows is not None:
        assert n_got == want_rows
    # pushdown is Inexact (table_provider.rs:941-962): it never drops a row the predicate keeps; the only extra rows are those
    # whose filtered column is NULL (record_filter.rs:107-110 lets them pass, DataFusion's own filter removes them later)
    key = lambda b: list(zip(b.column("name").to_pylist(), b.column("start").to_pylist(), b.column("flags").to_pylist(), b.column("end").to_pylist()))
    names_got = collections.Counter(k for b in got for k in key(b))
    names_manual = collections.Counter(zip(manual["name"].to_pylist(), manual["start"].to_pylist(), manual["flags"].to_pylist(), manual["end"].to_pylist()))
    assert not (names_manual - names_got), "pushdown dropped rows the predicate keeps"
    extra = names_got - names_manual
    assert all(k[3] is None for k in extra), "extra rows must be explained by a NULL in a filtered column"


def test_no_coor_only_file():
    # indexed_read_test.rs:244-271: 2 rows, chrom NULL, CB / CR present
    path = GOLDEN / "no_coor_only.bam"
    p = _provider(path, tag_fields=["CB", "CR"])
    plan = p.scan(None, [], None, target_partitions=4)
    t = plan.collect()
    assert t.num_rows == 2 and t["chrom"].null_count == 2 and t["start"].null_count == 2
    assert t["CB"].to_pylist() == ["CELL1", "CELL2"] and t["CR"].to_pylist() == ["RAW1", "RAW2"]


def test_unknown_chromosome_is_an_execution_error():
    import bamscan
    p = _provider(GOLDEN / "multi_chrom.bam")
    plan = p.scan(None, [("chrom", "=", ["chrNOPE"])], None, target_partitions=2)
    with pytest.raises(bamscan.BamScanError):
        plan.collect()


@pytest.mark.parametrize("tp", [1, 3, 8])
def test_synthetic_indexed_full_scan_with_unplaced_tail(syn_dir, tp):
    path = gen_bam(syn_dir, "short", 20000, seed=4, bai=True, unmapped=37)
    tags = ["NM", "MD"]
    p = _provider(path, tag_fields=tags, chunk_inflated_bytes=1 << 20)
    o = _oracle(path, tag_fields=tags)
    plan = p.scan(None, [], None, target_partitions=tp)
    n = plan.output_partition_count()
    assert plan.partition_ranges(n - 1)[0]["region_mode"] == 3            # the "*" partition comes last
    rows = 0
    for i in range(n):
        got = run_partition(plan, i)
        assert_same(got, expected_partition(o, plan, i, []), f"tp={tp} p{i}")
        rows += sum(b.num_rows for b in got)
    assert rows == 20037
    star = pa.Table.from_batches(run_partition(plan, n - 1))
    assert star.num_rows == 37 and star["chrom"].null_count == 37


@pytest.mark.parametrize("gpus", [1, 2, 4, 8])
def test_config4_region_query_partitioned(syn_dir, gpus):
    # BASELINE config 4: WHERE chrom='chr1' AND start BETWEEN x AND y, regions partitioned over 1/2/4/8 devices
    path = gen_bam(syn_dir, "short", 20000, seed=4, bai=True, unmapped=37)
    filters = [("chrom", "=", ["chr1"]), ("start", "between", [50_000_000, 150_000_000])]
    p = _provider(path, chunk_inflated_bytes=1 << 20)
    o = _oracle(path)
    plan = p.scan([0, 1, 2, 3, 6, 4, 9], filters, None, target_partitions=gpus)
    assert plan.output_partition_count() == gpus
    got_all = []
    for i in range(gpus):
        got = run_partition(plan, i)
        # an indexed range starts anywhere inside its first BGZF member (often several 16 KiB segments in): that exact start
        # must seed the record chain directly, not through the sequential repair path
        assert plan.last_stats["boundary_seam_mismatches"] == 0 and plan.last_stats["boundary_repairs"] == 0, plan.last_stats
        assert_same(got, expected_partition(o, plan, i, filters, projection=[0, 1, 2, 3, 6, 4, 9]), f"gpus={gpus} p{i}")
        got_all += got
    full = pa.Table.from_batches([o.scan(projection=[0, 1, 2, 3, 6, 4, 9])])
    manual = full.filter(pc.and_(pc.equal(full["chrom"], "chr1"), pc.and_(pc.greater_equal(full["start"], 50_000_000), pc.less_equal(full["start"], 150_000_000))))
    got_t = pa.Table.from_batches(got_all, schema=manual.schema)
    assert got_t.num_rows == manual.num_rows > 500
    assert got_t.equals(manual)       # sub-regions are emitted in genomic order, so the concatenation is the file-order subset


@pytest.mark.parametrize("depth", [5, 6])
def test_csi_index_scans_the_same_rows_as_the_bai(tmp_path, depth):
    """index_utils.rs:54-56, 68-76 discover `<path>.csi`; this build also reads it (CSIv1, any min_shift / depth): the rows of a
    region query planned from a CSI equal the rows planned from the BAI (indexed_read_large_test.rs pins 1662 for chr1)."""
    from conftest import bai_to_csi
    path = GOLDEN / "multi_chrom_large.bam"
    csi = bai_to_csi(str(path) + ".bai", tmp_path / "l.csi", depth=depth)
    for filters, want in (([("chrom", "=", ["chr1"])], 1662), ([("chrom", "=", ["chr2"]), ("start", "between", [1000, 20000])], None), ([], 4277)):
        tabs = []
        for prov in (_provider(path, tag_fields=["NM"]), _provider(path, tag_fields=["NM"], index_path=str(csi))):
            plan = prov.scan(None, filters, None, target_partitions=3)
            assert all(r["region_mode"] != 0 for i in range(plan.output_partition_count()) for r in plan.partition_ranges(i))
            tabs.append([b for i in range(plan.output_partition_count()) for b in run_partition(plan, i)])
        assert_same(tabs[1], tabs[0], f"csi depth {depth} {filters}")
        if want is not None:
            assert sum(b.num_rows for b in tabs[1]) == want
