"""CPU, world_size 2 over gloo: the N>1 path of bench.py / a multi-GPU deployment is 'one process per GPU, block-range
shard per rank, no data-path collective'.  Here each rank plans the same block-range partitions through the C ABI
(planning is host-only), takes its own partition, scans it with the ORACLE from an exact record start inside its first
block (the GPU finds that start by speculation), and the ranks only exchange row counts / checksums / timings --
exactly the control-plane traffic bench.py uses (barrier + all_reduce MAX)."""
import os
import socket
import sys

import pytest

from conftest import GOLDEN, ROOT, PKG, gen_bam


def _worker(rank, world, port, path, out_dir):
    for p in (str(ROOT), str(PKG), str(ROOT / "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import pyarrow as pa
    import bamscan
    from oracle.bam_oracle import OracleBam
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prov = bamscan.BamTableProvider(path)
    plan = prov.scan(None, [], None, target_partitions=world, partition_mode="block_range")
    assert plan.output_partition_count() == world
    (rng,) = plan.partition_ranges(rank)
    o = OracleBam(path)
    voff, _idx, _n = o.index_records(stride=0)             # first record start of every block (what speculation recovers on the GPU)
    def first_start_at_or_after(coff):
        cands = [int(v) for v in voff if (int(v) >> 16) >= coff]
        return cands[0] if cands else 0
    start = 0 if rng["exact_start"] else first_start_at_or_after(rng["coff_begin"])
    stop = first_start_at_or_after(rng["coff_end"]) if rng["stop_uoff"] != 2 ** 64 - 1 else 0
    if not rng["exact_start"] and start == 0:
        rows = 0; batch = None
    else:
        batch = o.scan(start_voffset=start, stop_voffset=stop)
        rows = batch.num_rows
    t = torch.tensor([rows], dtype=torch.int64)
    dist.all_reduce(t)                                      # control plane only
    tmax = torch.tensor([0.001 * (rank + 1)], dtype=torch.float64)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.barrier()
    if batch is not None:
        with pa.OSFile(os.path.join(out_dir, f"part{rank}.arrow"), "wb") as f, pa.ipc.new_file(f, batch.schema) as w:
            w.write_batch(batch)
    with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
        f.write(f"{int(t.item())} {float(tmax.item())}")
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["short", "long"])
def test_two_rank_block_range_shards(tmp_path, syn_dir, mode):
    import torch.multiprocessing as mp
    import pyarrow as pa
    from oracle.bam_oracle import OracleBam
    path = str(gen_bam(syn_dir, mode, 20000 if mode == "short" else 300, seed=7))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, path, str(tmp_path)), nprocs=2, join=True)
    full = OracleBam(path).scan()
    totals = [open(tmp_path / f"rank{r}.txt").read().split() for r in range(2)]
    assert all(int(t[0]) == full.num_rows for t in totals)          # every rank saw the same global row count
    assert all(abs(float(t[1]) - 0.002) < 1e-9 for t in totals)     # MAX over ranks
    parts = []
    for r in range(2):
        with pa.OSFile(str(tmp_path / f"part{r}.arrow"), "rb") as f:
            parts += pa.ipc.open_file(f).read_all().to_batches()
    assert pa.Table.from_batches(parts).equals(pa.Table.from_batches([full]))   # shards concatenate, in rank order, to the full scan
