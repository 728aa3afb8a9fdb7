"""CPU: the indexed-scan planning functions of the product, replaying the reference's own unit tests
(bio-format-core/src/genomic_filter.rs:350-557 and partition_balancer.rs:297-1005) through the C ABI mirrors
bamscan_extract_regions / bamscan_balance_partitions, plus plan-level pins of table_provider.rs:1001-1093."""
import pytest

from conftest import GOLDEN, gen_bam


def eg(filters, zero_based):
    import bamscan
    return bamscan.extract_genomic_regions(filters, zero_based)


def bp(estimates, target):
    import bamscan
    return bamscan.balance_partitions(estimates, target)


def est(chrom, b, contig_len=None, unmapped=0, bins=None, span=0):
    return dict(chrom=chrom, bytes=b, contig_length=contig_len, unmapped_count=unmapped, bins=bins, leaf_bin_span=span)


# ---------------------------------------------------------------- genomic_filter.rs tests
def test_extract_chrom_eq():
    a = eg([("chrom", "=", ["chr1"])], True)
    assert a["regions"] == [dict(chrom="chr1", start=None, end=None, unmapped_tail=False)] and a["residual"] == []


def test_extract_chrom_in_list():
    a = eg([("chrom", "in", ["chr2", "chr1", "chr1"])], True)
    assert [r["chrom"] for r in a["regions"]] == ["chr1", "chr2"]        # sorted + de-duplicated (:72-73)


@pytest.mark.parametrize("zero_based,lo,want_start", [(True, 999, 1000), (False, 1000, 1000)])
def test_extract_chrom_with_range(zero_based, lo, want_start):
    a = eg([("chrom", "=", ["chr1"]), ("start", ">=", [lo]), ("end", "<=", [2000])], zero_based)
    assert (a["regions"][0]["start"], a["regions"][0]["end"]) == (want_start, 2000) and a["residual"] == []


@pytest.mark.parametrize("zero_based,v", [(True, 999), (False, 1000)])
def test_exact_start_bounds_region(zero_based, v):
    a = eg([("chrom", "=", ["chr1"]), ("start", "=", [v])], zero_based)
    assert (a["regions"][0]["start"], a["regions"][0]["end"]) == (1000, 1000)


def test_contradictory_start_bounds_unsatisfiable():
    a = eg([("chrom", "=", ["chr1"]), ("start", "=", [1000]), ("start", ">", [1000])], False)
    assert a["unsatisfiable"] and a["regions"] == [] and a["residual"] == []


@pytest.mark.parametrize("zero_based,lo,hi", [(False, 1000, 2000), (True, 999, 1999)])
def test_start_upper_bound(zero_based, lo, hi):
    a = eg([("chrom", "=", ["chr1"]), ("start", ">=", [lo]), ("start", "<=", [hi])], zero_based)
    assert (a["regions"][0]["start"], a["regions"][0]["end"]) == (1000, 2000)


def test_start_exclusive_upper_bound_one_based():
    a = eg([("chrom", "=", ["chr1"]), ("start", ">=", [1000]), ("start", "<", [2000])], False)
    assert (a["regions"][0]["start"], a["regions"][0]["end"]) == (1000, 1999)


def test_non_genomic_filter_becomes_residual():
    a = eg([("chrom", "=", ["chr1"]), ("mapping_quality", ">=", [30])], True)
    assert len(a["regions"]) == 1 and a["residual"] == [1]
    b = eg([("mapping_quality", ">=", [30])], True)
    assert b["regions"] == [] and not b["unsatisfiable"] and b["residual"] == [0]


def test_between_start_with_and_without_chrom():
    assert eg([("start", "between", [999, 1999])], True)["regions"] == []
    a = eg([("chrom", "=", ["chr1"]), ("start", "between", [999, 1999])], True)
    assert (a["regions"][0]["start"], a["regions"][0]["end"]) == (1000, 2000)


def test_shapes_that_stay_residual():
    # chrom != , start != , end > , NOT BETWEEN, NOT IN: genomic_filter.rs:188-189,225-227,259-261,299,324-327
    a = eg([("chrom", "=", ["chr1"]), ("chrom", "!=", ["chr2"]), ("start", "!=", [5]), ("end", ">", [7]), ("start", "not_between", [1, 2]),
            ("chrom", "not_in", ["chrX"]), ("end", "<", [100]), ("start", ">=", [2.5])], False)
    assert a["residual"] == [1, 2, 3, 4, 5, 7] and a["end"] == 99 and a["start"] is None


# ---------------------------------------------------------------- partition_balancer.rs tests
def test_empty_and_single_partition():
    assert bp([], 4) == []
    r = bp([est("chr1", 100), est("chr2", 50), est("chrX", 30)], 1)
    assert len(r) == 1 and len(r[0]["regions"]) == 3 and r[0]["total_estimated_bytes"] == 180


def test_uniform_and_skewed():
    r = bp([est(c, 100) for c in ("chr1", "chr2", "chr3", "chr4")], 2)
    assert [p["total_estimated_bytes"] for p in r] == [200, 200]
    r = bp([est("chr1", 100), est("chr2", 50), est("chr3", 10)], 2)                      # unsplittable
    assert [p["total_estimated_bytes"] for p in r] == [100, 60]
    r = bp([est("chr1", 100, 249_000_000), est("chr2", 50), est("chr3", 10)], 2)        # splittable at the byte budget
    assert [p["total_estimated_bytes"] for p in r] == [80, 80]
    assert r[0]["regions"] == [dict(chrom="chr1", start=1, end=199_200_000, unmapped_tail=False)]
    assert r[1]["regions"][0] == dict(chrom="chr1", start=199_200_001, end=None, unmapped_tail=False)


def test_region_splitting_and_counts():
    r = bp([est("chr1", 200, 249_000_000), est("chr2", 10)], 4)
    assert sum(len(p["regions"]) for p in r) > 2 and len(r) <= 4
    r = bp([est(f"chr{i}", 0) for i in range(1, 5)], 2)                                   # all-zero: round robin
    assert [len(p["regions"]) for p in r] == [2, 2]
    assert len(bp([est("chr1", 100), est("chr2", 50)], 8)) == 2
    r = bp([est("chr1", 100, 249_000_000), est("chr2", 50, 243_000_000)], 8)
    assert 2 < len(r) <= 8
    r = bp([est("chr1", 100)], 4)
    assert len(r) == 1 and len(r[0]["regions"]) == 1


HUMAN = [("chr1", 249), ("chr2", 243), ("chr3", 198), ("chr4", 191), ("chr5", 181), ("chr6", 171), ("chr7", 159), ("chr8", 146),
         ("chr9", 141), ("chr10", 136), ("chr11", 135), ("chr12", 134), ("chr13", 115), ("chr14", 107), ("chr15", 102), ("chr16", 90),
         ("chr17", 84), ("chr18", 80), ("chr19", 59), ("chr20", 64), ("chr21", 47), ("chr22", 51), ("chrX", 155), ("chrY", 57)]


def test_never_exceeds_target_and_keeps_all_regions():
    r = bp([est(c, b, l) for c, b, l in [("chr1", 100, 249_000_000), ("chr2", 90, 243_000_000), ("chr3", 80, 198_000_000),
                                        ("chr4", 70, 191_000_000), ("chr5", 60, 181_000_000), ("chrX", 50, 155_000_000)]], 4)
    assert len(r) <= 4
    r = bp([est("chr1", 100), est("chr2", 50), est("chr3", 30), est("chrX", 20)], 3)
    assert sorted({g["chrom"] for p in r for g in p["regions"]}) == ["chr1", "chr2", "chr3", "chrX"]


@pytest.mark.parametrize("target", [2, 4, 8])
def test_single_contig_splits_to_target_partitions(target):
    r = bp([est("chr1", 1000, 249_000_000)], target)
    assert len(r) == target
    regs = [g for p in r for g in p["regions"]]
    assert all(g["chrom"] == "chr1" and g["start"] is not None for g in regs)
    assert all(g["end"] is not None for g in regs[:-1]) and regs[-1]["end"] is None      # last piece open-ended


def test_realistic_human_genome_distribution():
    r = bp([est(c, b, b * 1_000_000) for c, b in HUMAN], 8)
    assert 0 < len(r) <= 8
    tot = [p["total_estimated_bytes"] for p in r]
    assert sum(tot) == sum(b for _, b in HUMAN) and max(tot) <= 2 * min(tot)


def test_unmapped_tail_emission():
    r = bp([est("chr1", 200, 249_000_000, unmapped=1000), est("chr2", 10)], 4)
    tails = [g for p in r for g in p["regions"] if g["unmapped_tail"]]
    assert tails == [dict(chrom="chr1", start=None, end=None, unmapped_tail=True)]
    r = bp([est("chr1", 100, 249_000_000, 500), est("chr2", 95, 243_000_000, 300), est("chrM", 5, 16_569, 100)], 4)
    assert len([g for p in r for g in p["regions"] if g["chrom"] == "chrM" and g["unmapped_tail"]]) == 1


def test_linear_scan_perfect_balance():
    r = bp([est("chr1", 249, 249_000_000), est("chr2", 243, 243_000_000), est("chr3", 198, 198_000_000), est("chrX", 60, 155_000_000)], 4)
    assert len(r) == 4 and sum(p["total_estimated_bytes"] for p in r) == 750
    assert all(187 <= p["total_estimated_bytes"] <= 189 for p in r)


def test_zero_byte_regions_not_dropped():
    r = bp([est("chr1", 0), est("chr2", 100), est("chrX", 0)], 4)
    assert sorted(g["chrom"] for p in r for g in p["regions"]) == ["chr1", "chr2", "chrX"]


@pytest.mark.parametrize("target", [2, 3, 4, 8, 16])
def test_total_bytes_preserved(target):
    r = bp([est("chr1", 249, 249_000_000), est("chr2", 243, 243_000_000), est("chr3", 198, 198_000_000)], target)
    assert sum(p["total_estimated_bytes"] for p in r) == 690


def test_bin_aware_split_concentrates_on_data():
    span = 16384
    r = bp([est("chr1", 1000, 249_000_000, bins=[i * span + 1 for i in range(100)], span=span)], 4)
    assert len(r) == 4
    regs = [g for p in r for g in p["regions"]]
    assert all(g["end"] < 2_000_000 for g in regs[:-1] if g["end"] is not None)


@pytest.mark.parametrize("target", [2, 4, 8])
def test_bin_aware_preserves_total(target):
    span = 16384
    r = bp([est("chr1", 500, 249_000_000, bins=[i * span + 1 for i in range(200)], span=span)], target)
    assert sum(p["total_estimated_bytes"] for p in r) == 500


def test_bin_aware_fallback_when_no_bins():
    a = bp([est("chr1", 200, 249_000_000, bins=[], span=0)], 4)
    b = bp([est("chr1", 200, 249_000_000)], 4)
    assert a == b


def test_bin_aware_sparse_wes_pattern():
    span = 16384
    bins = sorted([10_000_000 + i * span for i in range(40)] + [50_000_000 + i * span for i in range(30)] + [200_000_000 + i * span for i in range(30)])
    r = bp([est("chr1", 600, 249_000_000, bins=bins, span=span)], 4)
    assert len(r) == 4 and sum(p["total_estimated_bytes"] for p in r) == 600
    for e in [g["end"] for p in r for g in p["regions"] if g["end"] is not None]:
        assert 10_000_000 <= e <= 11_000_000 or 50_000_000 <= e <= 51_000_000 or 200_000_000 <= e <= 201_000_000


def test_zero_estimate_regions_distributed_evenly():
    ests = [est(c, b, b * 1_000_000) for c, b in HUMAN] + [est(f"alt_contig_{i}", 0) for i in range(60)]
    r = bp(ests, 8)
    assert sum(len(p["regions"]) for p in r) >= 84 and sum(p["total_estimated_bytes"] for p in r) == 3095
    z = [sum(1 for g in p["regions"] if g["chrom"].startswith("alt_contig_")) for p in r]
    assert max(z) <= 15 and min(z) >= 1
    ests = [est("chr1", 100, 249_000_000), est("chr2", 80, 243_000_000), est("chr3", 60, 198_000_000)] + [est(f"scaffold_{i}", 0) for i in range(20)]
    r = bp(ests, 4)
    s = [sum(1 for g in p["regions"] if g["chrom"].startswith("scaffold_")) for p in r]
    assert sum(p["total_estimated_bytes"] for p in r) == 240 and max(s) <= 8


# ---------------------------------------------------------------- plan-level pins (table_provider.rs:1001-1093)
def _provider(path, **kw):
    import bamscan
    return bamscan.BamTableProvider(str(path), None, True, kw.pop("tag_fields", None), False, True, 100, None, **kw)


@pytest.mark.parametrize("tp", [1, 2, 3, 4, 8])
def test_indexed_full_scan_partition_count_never_exceeds_target(tp):
    p = _provider(GOLDEN / "multi_chrom.bam")
    plan = p.scan(None, [], None, target_partitions=tp)
    n = plan.output_partition_count()
    assert 1 <= n <= tp
    regs = [g for i in range(n) for g in plan.partition_regions(i)]
    assert {g["ref"] for g in regs} == {0, 1, 2}                       # one (possibly split) region per reference
    # BAI metadata (mapped, unmapped) = (159,1) (157,2) (100,2) => a tail per reference; the target == 1 fast path emits none
    # (partition_balancer.rs:71-79)
    assert sum(1 for g in regs if g["unmapped_tail"]) == (3 if tp > 1 else 0)


def test_unsatisfiable_filters_give_empty_exec():
    p = _provider(GOLDEN / "multi_chrom.bam")
    plan = p.scan(None, [("chrom", "=", ["chr1"]), ("start", ">", [1000]), ("start", "<", [500])], None, target_partitions=4)
    assert plan.output_partition_count() == 0                          # EmptyExec (table_provider.rs:1005-1010)


def test_no_coor_file_gets_star_partition():
    # table_provider.rs:1048-1056, :1180-1206 ; tests/no_coor_only.bam(.bai): n_no_coor = 2, one reference without bins
    p = _provider(GOLDEN / "no_coor_only.bam")
    plan = p.scan(None, [], None, target_partitions=4)
    n = plan.output_partition_count()
    last = plan.partition_regions(n - 1)
    assert last == [dict(ref=-1, start=None, end=None, unmapped_tail=True, estimated_bytes=2)]
    assert [r["region_mode"] for r in plan.partition_ranges(n - 1)] == [3]


def test_region_query_ranges_cover_only_needed_blocks(syn_dir):
    path = gen_bam(syn_dir, "short", 20000, seed=4, bai=True)
    p = _provider(path)
    total_blocks = p.scan(None, [], None, partition_mode="block_range").partition_ranges(0)[0]["block_end"]
    plan = p.scan(None, [("chrom", "=", ["chr2"]), ("start", "between", [100_000_000, 120_000_000])], None, target_partitions=2)
    ranges = [r for i in range(plan.output_partition_count()) for r in plan.partition_ranges(i)]
    assert ranges and all(r["region_mode"] in (1, 2) and r["region_ref"] == 1 for r in ranges)
    touched = sum(r["block_end"] - r["block_begin"] for r in ranges if r["region_mode"] == 1)
    assert touched < total_blocks // 4                                   # the index pruned most of the file


# ---------------------------------------------------------------- CSI (index_utils.rs:54-56, 68-76; CSIv1)
def _plan_ranges(p, filters, tp):
    plan = p.scan(None, filters, None, target_partitions=tp)
    n = plan.output_partition_count()
    return [[(r["region_mode"], r["region_ref"], r["block_begin"], r["block_end"]) for r in plan.partition_ranges(i)] for i in range(n)], \
           [plan.partition_regions(i) for i in range(n)]


CSI_QUERIES = [[], [("chrom", "=", ["chr1"])], [("chrom", "=", ["chr2"]), ("start", "between", [100_000_000, 120_000_000])],
               [("chrom", "in", ["chr1", "chr3"]), ("start", ">", [30_000_000])], [("chrom", "=", ["chr2"]), ("start", "<", [5_000_000])]]


@pytest.mark.parametrize("depth,bgzf", [(5, True), (5, False), (6, True), (9, True)])
def test_csi_plans_like_the_bai(syn_dir, tmp_path, depth, bgzf):
    """A CSI holding the BAI's own bins (same scheme, or one level deeper) must plan the same partitions and regions; a range may
    only start EARLIER than the BAI's (loffset of an ancestor bin instead of the 16 KiB linear-index window), never later."""
    from conftest import bai_to_csi
    path = gen_bam(syn_dir, "short", 20000, seed=4, bai=True)
    csi = bai_to_csi(str(path) + ".bai", tmp_path / "x.csi", depth=depth, bgzf=bgzf)
    import bamscan
    pb = _provider(path)
    pc = bamscan.BamTableProvider(str(path), None, True, None, False, True, 100, None, index_path=str(csi))
    for filters in CSI_QUERIES:
        for tp in (1, 4):
            rb, gb = _plan_ranges(pb, filters, tp)
            rc, gc = _plan_ranges(pc, filters, tp)
            assert gb == gc, (filters, tp)                                   # estimates, balancer input and output are the same
            assert len(rb) == len(rc)
            for qb, qc in zip(rb, rc):
                assert [(m, r) for m, r, _, _ in qb if m != 1] == [(m, r) for m, r, _, _ in qc if m != 1]
                beg_b = min((b for m, _, b, _ in qb if m == 1), default=None); beg_c = min((b for m, _, b, _ in qc if m == 1), default=None)
                end_b = max((e for m, _, _, e in qb if m == 1), default=None); end_c = max((e for m, _, _, e in qc if m == 1), default=None)
                assert (beg_b is None) == (beg_c is None)
                if beg_b is not None:
                    assert beg_c <= beg_b and end_c == end_b, (filters, tp, qb, qc)


def test_csi_is_discovered_after_bai(syn_dir, tmp_path):
    """discover_bam_index: `<path>.bai`, `<stem>.bai`, then `<path>.csi`."""
    import shutil
    from conftest import bai_to_csi
    src = gen_bam(syn_dir, "short", 20000, seed=4, bai=True)
    bam = tmp_path / "only_csi.bam"
    shutil.copy(src, bam)
    assert _provider(bam).scan(None, [("chrom", "=", ["chr1"])], None, target_partitions=2).partition_ranges(0)[0]["region_mode"] == 0   # no index: full scan
    bai_to_csi(str(src) + ".bai", str(bam) + ".csi")
    plan = _provider(bam).scan(None, [("chrom", "=", ["chr1"])], None, target_partitions=2)
    assert all(r["region_mode"] in (1, 2) for i in range(plan.output_partition_count()) for r in plan.partition_ranges(i))


def test_csi_fixture_metadata_and_no_coor(tmp_path):
    from conftest import bai_to_csi
    import bamscan
    csi = bai_to_csi(GOLDEN / "no_coor_only.bam.bai", tmp_path / "n.csi", depth=6)
    p = bamscan.BamTableProvider(str(GOLDEN / "no_coor_only.bam"), None, True, None, False, True, 100, None, index_path=str(csi))
    plan = p.scan(None, [], None, target_partitions=4)
    assert plan.partition_regions(plan.output_partition_count() - 1) == [dict(ref=-1, start=None, end=None, unmapped_tail=True, estimated_bytes=2)]
    csi = bai_to_csi(GOLDEN / "multi_chrom.bam.bai", tmp_path / "m.csi", depth=6)
    p = bamscan.BamTableProvider(str(GOLDEN / "multi_chrom.bam"), None, True, None, False, True, 100, None, index_path=str(csi))
    plan = p.scan(None, [], None, target_partitions=4)
    regs = [g for i in range(plan.output_partition_count()) for g in plan.partition_regions(i)]
    assert sum(1 for g in regs if g["unmapped_tail"]) == 3               # the metadata pseudo-bin (first_bin(depth + 1) + 1) was read


def test_corrupt_csi_is_refused(tmp_path):
    import bamscan
    bad = tmp_path / "bad.csi"
    bad.write_bytes(b"CSI\x01" + b"\x0e\0\0\0\x05\0\0\0\0\0\0\0" + b"\xff\xff\xff\x7f")           # n_ref the file cannot hold
    with pytest.raises(Exception):
        bamscan.BamTableProvider(str(GOLDEN / "multi_chrom.bam"), None, True, None, False, True, 100, None, index_path=str(bad))
    bad.write_bytes(b"CSI\x01" + b"\x0e\0\0\0\x40\0\0\0\0\0\0\0\0\0\0\0")                        # depth 64
    with pytest.raises(Exception):
        bamscan.BamTableProvider(str(GOLDEN / "multi_chrom.bam"), None, True, None, False, True, 100, None, index_path=str(bad))


@pytest.mark.parametrize("depth", [5, 6])
def test_csi_planned_ranges_hold_the_pinned_rows(tmp_path, depth):
    """CPU: the oracle run over the ranges a CSI plan exposes returns the rows the reference pins for the BAI
    (indexed_read_large_test.rs:63,85,95: 1662 / 1694 / 921) -- the GPU twin is test_csi_index_scans_the_same_rows_as_the_bai."""
    from conftest import bai_to_csi
    from oracle.bam_oracle import OracleBam
    import bamscan
    path = GOLDEN / "multi_chrom_large.bam"
    csi = bai_to_csi(str(path) + ".bai", tmp_path / "l.csi", depth=depth)
    p = bamscan.BamTableProvider(str(path), None, True, None, False, True, 100, None, index_path=str(csi))
    o = OracleBam(str(path))
    for chrom, want in (("chr1", 1662), ("chr2", 1694), ("chrX", 921)):
        plan = p.scan(None, [("chrom", "=", [chrom])], None, target_partitions=3)
        rows = 0
        for i in range(plan.output_partition_count()):
            for r in plan.partition_ranges(i):
                rows += o.scan(projection=[0], start_voffset=r["start_voffset"], stop_voffset=r["stop_voffset"],
                               region=(r["region_mode"], r["region_ref"], r["region_start"], r["region_end"]), filters=[]).num_rows
        assert rows == want, (chrom, depth)


def test_corrupt_indices_never_crash_the_planner(tmp_path):
    """Truncated / bit-flipped BAI and CSI files either plan or raise BamScanError (an index is untrusted input; a count the file
    cannot hold must not become an allocation)."""
    import random
    from conftest import bai_to_csi
    import bamscan
    bam = GOLDEN / "multi_chrom_large.bam"
    bai = (GOLDEN / "multi_chrom_large.bam.bai").read_bytes()
    csi = bai_to_csi(GOLDEN / "multi_chrom_large.bam.bai", tmp_path / "a.csi", depth=6, bgzf=False).read_bytes()
    rng, planned = random.Random(7), 0
    for it in range(400):
        b = bytearray(bai if it % 2 == 0 else csi)
        mode = rng.randrange(3)
        if mode == 0:
            b = b[:rng.randrange(len(b))]
        elif mode == 1:
            for _ in range(rng.randrange(1, 6)):
                b[rng.randrange(len(b))] = rng.randrange(256)
        else:
            pos = rng.randrange(0, 200)
            b[pos:pos + 4] = rng.randrange(2 ** 32).to_bytes(4, "little")
        p = tmp_path / "f.idx"
        p.write_bytes(bytes(b))
        try:
            pr = bamscan.BamTableProvider(str(bam), None, True, None, False, True, 100, None, index_path=str(p))
            for f in ([], [("chrom", "=", ["chr1"]), ("start", "between", [1000, 50000])]):
                plan = pr.scan(None, f, None, target_partitions=4)
                for i in range(plan.output_partition_count()):
                    plan.partition_ranges(i)
            planned += 1
        except bamscan.BamScanError:
            pass
    assert planned > 0


@pytest.mark.parametrize("index", ["bai", "csi6"])
def test_random_region_plans_lose_no_row(syn_dir, tmp_path, index):
    """Completeness of the index query incl. this build's upper-bound chunk prune (host_index.cpp::bai_query): for random regions
    the oracle, run over nothing but the ranges the plan exposes, finds exactly the rows a brute-force filter of the whole file
    finds (row rule: chrom == c and start in [lo, hi], physical_exec.rs:1295-1314)."""
    import random
    import pyarrow.compute as pc
    from conftest import bai_to_csi
    from oracle.bam_oracle import OracleBam
    import bamscan
    path = gen_bam(syn_dir, "short", 20000, seed=4, bai=True)
    kw = {}
    if index == "csi6":
        kw["index_path"] = str(bai_to_csi(str(path) + ".bai", tmp_path / "r.csi", depth=6))
    p = bamscan.BamTableProvider(str(path), None, True, None, False, True, 100, None, **kw)
    o = OracleBam(str(path))
    full = o.scan(projection=[1, 2])
    chrom, start = full.column(0), full.column(1)
    names = [c for c in pc.unique(chrom).to_pylist() if c is not None]
    rng = random.Random(11)
    checked = 0
    for it in range(24):
        c = rng.choice(names)
        starts = pc.filter(start, pc.equal(chrom, c)).to_pylist()
        lo = rng.choice(starts) + rng.randrange(-5000, 5000)
        hi = lo + rng.choice([0, 1000, 100_000, 5_000_000, 80_000_000])
        lo = max(lo, 0)
        filters = [("chrom", "=", [c]), ("start", "between", [lo, hi])] if it % 3 else [("chrom", "=", [c]), ("start", ">=", [lo])]
        if it % 3 == 0:
            hi = 1 << 40
        want = sum(1 for s in starts if s is not None and lo <= s <= hi)
        plan = p.scan(None, filters, None, target_partitions=rng.choice([1, 3, 8]))
        got = 0
        for i in range(plan.output_partition_count()):
            for r in plan.partition_ranges(i):
                if r["region_mode"] != 1:
                    continue                                             # per-reference unmapped tails: start is NULL, never in a start range
                got += o.scan(projection=[0], start_voffset=r["start_voffset"], stop_voffset=r["stop_voffset"],
                              region=(r["region_mode"], r["region_ref"], r["region_start"], r["region_end"]),
                              filters=[f for f in filters if f[0] == "start"]).num_rows
        assert got == want, (filters, got, want)
        checked += want
    assert checked > 0


def test_balance_partitions_tiles_every_contig_property():
    """Property (hypothesis): whatever the estimates, the sub-regions of a contig tile it -- first piece starts at 1 (or is the
    whole contig), each next piece starts one base after the previous end, only the last piece is open-ended -- so the row rule
    `start in [region.start, region.end]` (physical_exec.rs:1295-1314) can neither duplicate nor lose a row across partitions;
    the partition count never exceeds the target and the byte estimates are conserved (partition_balancer.rs:61-295)."""
    from hypothesis import given, settings, strategies as st
    import bamscan

    contig = st.tuples(st.integers(0, 10 ** 10), st.one_of(st.none(), st.integers(1, 3 * 10 ** 8)), st.integers(0, 5),
                       st.one_of(st.none(), st.lists(st.integers(0, 2000), min_size=1, max_size=40, unique=True)))

    @settings(max_examples=300, deadline=None)
    @given(st.lists(contig, min_size=1, max_size=8), st.integers(1, 24))
    def check(contigs, target):
        ests = []
        for i, (b, length, unm, bins) in enumerate(contigs):
            pos = sorted(k * 16384 + 1 for k in bins) if bins else None
            if pos and length:
                pos = [x for x in pos if x <= length] or None
            ests.append(est(f"c{i}", b, contig_len=length, unmapped=unm, bins=pos, span=16384 if pos else 0))
        parts = bamscan.balance_partitions(ests, target)
        assert len(parts) <= target
        total = sum(e["bytes"] for e in ests)
        n_tail_regions = sum(1 for p in parts for g in p["regions"] if g["unmapped_tail"])
        assert sum(p["total_estimated_bytes"] for p in parts) == total + n_tail_regions      # a tail counts one byte (:267)
        pieces, tails = {}, {}
        for p in parts:
            for g in p["regions"]:
                (tails if g["unmapped_tail"] else pieces).setdefault(g["chrom"], []).append(g)
        assert set(pieces) == {e["chrom"] for e in ests}
        for e in ests:
            ps = pieces[e["chrom"]]                                        # in partition order = genomic order
            if len(ps) == 1:
                assert ps[0]["start"] in (None, 1) and ps[0]["end"] is None
            else:
                assert ps[0]["start"] == 1 and ps[-1]["end"] is None
                for a, b in zip(ps, ps[1:]):
                    assert a["end"] is not None and a["end"] >= a["start"] and b["start"] == a["end"] + 1
            n_tails = len(tails.get(e["chrom"], []))
            # tails: none on the target == 1 and all-zero fast paths (:71-95), else one per reference with unmapped reads
            if target == 1 or total == 0:
                assert n_tails == 0
            elif e["bytes"] > 0:
                assert n_tails == (1 if e["unmapped_count"] > 0 else 0)
            else:
                assert n_tails <= 1

    check()


def test_extract_regions_is_a_superset_property():
    """Property (hypothesis): the pushed-down region is sound -- every position whose `start` passes all the conjuncts lies inside
    [region.start, region.end] (1-based, after the coordinate shift), and `unsatisfiable` is only claimed when no position passes
    (genomic_filter.rs:51-329: the pushdown is Inexact, it may keep too much, never too little)."""
    from hypothesis import given, settings, strategies as st

    val = st.integers(0, 400)
    conj = st.one_of(st.tuples(st.sampled_from([">", ">=", "<", "<=", "="]), st.tuples(val)),
                     st.tuples(st.just("between"), st.tuples(val, val)))

    def passes(pos, op, v):
        return {">": lambda: pos > v[0], ">=": lambda: pos >= v[0], "<": lambda: pos < v[0], "<=": lambda: pos <= v[0],
                "=": lambda: pos == v[0], "between": lambda: v[0] <= pos <= v[1]}[op]()

    @settings(max_examples=400, deadline=None)
    @given(st.lists(conj, min_size=0, max_size=4), st.booleans())
    def check(conjs, zero_based):
        filters = [("chrom", "=", ["chr1"])] + [("start", op, list(v)) for op, v in conjs]
        r = eg(filters, zero_based)
        ok = [p for p in range(0 if zero_based else 1, 402) if all(passes(p, op, v) for op, v in conjs)]
        if r["unsatisfiable"]:
            assert not ok, (filters, ok[:3])
            return
        assert [g["chrom"] for g in r["regions"]] == ["chr1"]
        lo, hi = r["regions"][0]["start"], r["regions"][0]["end"]
        for p in ok:
            one = p + 1 if zero_based else p
            assert (lo is None or one >= lo) and (hi is None or one <= hi), (filters, p, lo, hi)

    check()
