#!/usr/bin/env python3
"""Host<->device copy bandwidth with all ranks copying at the same time (pinned memory, 1 GiB per direction per rank):
tells how much of the N-GPU end-to-end number is the box's PCIe / host-memory ceiling.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_all.py"""
import os, sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "datafusion-bio-formats_b200"))
import torch, torch.distributed as dist
import bamscan
rank = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
bamscan.probe_pcie(rank, 1 << 28)                      # warm-up: context, pinned allocation paths
for rep in range(2):
    if world > 1:
        dist.barrier()
    r = bamscan.probe_pcie(rank, 1 << 30)
    vals = torch.tensor([r["h2d_gbps"], r["d2h_gbps"], r["bidir_gbps"]], dtype=torch.float64, device="cuda")
    if world > 1:
        allv = [torch.zeros_like(vals) for _ in range(world)]
        dist.all_gather(allv, vals)
    else:
        allv = [vals]
    if rank == 0:
        m = torch.stack(allv).cpu()
        print(json.dumps({"ranks": world, "rep": rep, "h2d_gbps_per_rank_min_mean": [float(m[:, 0].min()), float(m[:, 0].mean())],
                          "d2h_gbps_per_rank_min_mean": [float(m[:, 1].min()), float(m[:, 1].mean())],
                          "bidir_gbps_per_rank_min_mean": [float(m[:, 2].min()), float(m[:, 2].mean())],
                          "aggregate_d2h_gbps": float(m[:, 1].sum()), "aggregate_bidir_gbps": float(m[:, 2].sum())}), flush=True)
if world > 1:
    dist.destroy_process_group()
