#!/usr/bin/env python3
"""Write-path measurement: scan a synthetic config-2 BAM on the GPU, write the batches back through
bamscan_writer_*, report device and end-to-end rates, the compression ratio against the zlib-6 original, and the CPU cost
of the same job (zlib level 6 on one thread over a bounded sample of the stream)."""
import json
import os
import sys
import time
import zlib
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "datafusion-bio-formats_b200")):
    sys.path.insert(0, p)

import pyarrow as pa  # noqa: E402

import bench  # noqa: E402


def main():
    reads = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    mode = sys.argv[2] if len(sys.argv) > 2 else "short"
    import bamscan
    path, info = bench.ensure_bam(reads, 2 if mode == "short" else 5, False, mode=mode)
    tags = ["NM", "MD", "AS", "RG"] if mode == "short" else ["NM", "MD", "MM", "ML"]
    p = bamscan.BamTableProvider(str(path), None, True, tags, index_path="")
    batches = list(p.scan(None, [], None).execute(0))
    rows = sum(b.num_rows for b in batches)
    out = Path(os.environ.get("BAMSCAN_BENCH_DIR", "/dev/shm/bamscan_bench")) / "written.bam"
    res = {}
    for rep in range(3):
        ex = bamscan.BamWriteExec(str(out), p.schema(), tags, True, {"bio.bam.sort_order": "unsorted"})
        print(f"-- rep {rep}", file=sys.stderr, flush=True)
        t0 = time.perf_counter()
        n = ex.execute(batches)
        dt = time.perf_counter() - t0
        st = ex.stats
        res = {"rows": n, "e2e_s": dt, "e2e_reads_per_s": n / dt, "device_ms": st["ms_total"], "encode_ms": st["ms_encode"], "deflate_ms": st["ms_deflate"],
               "device_reads_per_s": n / (st["ms_total"] / 1e3), "bam_bytes": st["bam_bytes"], "arrow_bytes": st["arrow_bytes"], "file_bytes": st["compressed_bytes"],
               "deflate_gbps_uncompressed": st["bam_bytes"] / (st["ms_deflate"] / 1e3) / 1e9, "encode_gbps": (st["bam_bytes"] + st["arrow_bytes"]) / (st["ms_encode"] / 1e3) / 1e9,
               "members": st["members"], "kernel_launches": st["kernel_launches"]}
    assert res["rows"] == rows
    if len(sys.argv) > 3 and sys.argv[3] == "device":
        # device-resident transcode: the scan's batches stay in HBM (bamscan_next_device) and are written in place
        for rep in range(3):
            ex = bamscan.BamWriteExec(str(out), p.schema(), tags, True, {"bio.bam.sort_order": "unsorted"})
            t0 = time.perf_counter()
            n = ex.execute(p.scan(None, [], None).execute_device(0))
            dt = time.perf_counter() - t0
            res["transcode_device"] = {"rows": n, "e2e_s": dt, "reads_per_s": n / dt, "write_device_ms": ex.stats["ms_total"],
                                       "note": "scan (H2D of the compressed file, inflate, decode) -> device batches -> encode + deflate -> D2H of the compressed members -> file"}
    res["zlib6_file_bytes"] = info["compressed_bytes"]
    res["size_vs_zlib6"] = res["file_bytes"] / info["compressed_bytes"]
    res["ratio"] = res["bam_bytes"] / res["file_bytes"]
    # read back: the written file scans to the same rows
    p2 = bamscan.BamTableProvider(str(out), None, True, tags, index_path="")
    back = sum(b.num_rows for b in p2.scan(None, [], None).execute(0))
    res["read_back_rows"] = back
    # CPU: zlib level 6 + crc32 of 64 MiB of the same stream, one thread
    raw = b""
    data = open(path, "rb").read(40 << 20)
    off = 0
    import struct
    chunks = []
    while off + 28 <= len(data) and sum(map(len, chunks)) < (64 << 20):
        bsize = struct.unpack_from("<H", data, off + 16)[0] + 1
        if off + bsize > len(data):
            break
        chunks.append(zlib.decompress(data[off + 18: off + bsize - 8], -15)); off += bsize
    t0 = time.perf_counter()
    nb = 0
    for c in chunks:
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        co.compress(c); co.flush(); zlib.crc32(c); nb += len(c)
    dt = time.perf_counter() - t0
    res["cpu_zlib6_gbps_one_thread"] = nb / dt / 1e9
    res["mode"] = mode; res["reads"] = reads
    print(json.dumps(res))


if __name__ == "__main__":
    main()
