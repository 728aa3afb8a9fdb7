#!/usr/bin/env python3
"""BGZF-FASTQ scan on one GPU: device-resident CUDA-event time and end-to-end wall time for 1 and 8 block-range
partitions, the inflate launch's roofline fraction, and the CPU restatement (oracle/fastq_oracle.py, one thread) on a sample.
The workload mirrors the reference's own adjacent-path number (26.5 M reads of 101 bp, BGZF; SURVEY 5: 17.84 s on one core,
2.37 s on eight): usage  python tools/measure_fastq.py [reads=20000000]  -> markdown on stdout."""
import json, os, struct, sys, time, zlib
from concurrent.futures import ProcessPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "datafusion-bio-formats_b200"))
reads = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
READ_LEN = 101


def make_text(i0, n, seed):
    """n fixed-width records '@SIM.<9 digits> HSQ:<9 digits>/1' / 101 bases / '+' / 101 qualities, as bytes (vectorised)."""
    rng = np.random.default_rng(seed)
    head = b"@SIM.000000000 HSQ:000000000/1\n"
    W = len(head) + READ_LEN + 3 + READ_LEN + 1
    a = np.empty((n, W), np.uint8)
    a[:, :len(head)] = np.frombuffer(head, np.uint8)
    idx = np.arange(i0, i0 + n, dtype=np.int64)
    for d in range(9):
        dig = (idx // 10 ** (8 - d)) % 10 + 48
        a[:, 5 + d] = dig
        a[:, 19 + d] = (dig - 48 + d) % 10 + 48
    p = len(head)
    a[:, p:p + READ_LEN] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, (n, READ_LEN))]
    p += READ_LEN
    a[:, p:p + 3] = np.frombuffer(b"\n+\n", np.uint8)
    p += 3
    a[:, p:p + READ_LEN] = rng.integers(35, 75, (n, READ_LEN), dtype=np.uint8)
    a[:, p + READ_LEN] = 10
    return a.tobytes()


def bgzf_blocks(args):
    i0, n, seed = args
    data = make_text(i0, n, seed)
    out = []
    for i in range(0, len(data), 0xff00):
        chunk = data[i:i + 0xff00]
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = c.compress(chunk) + c.flush()
        out.append(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(body) + 25) + body + struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk)))
    return b"".join(out), len(data)


def ensure_file():
    base = Path(os.environ.get("BAMSCAN_BENCH_DIR", "/dev/shm/bamscan_bench")); base.mkdir(parents=True, exist_ok=True)
    path = base / f"sim_{reads}.fastq.bgz"
    meta = path.with_suffix(".json")
    if path.exists() and meta.exists():
        return path, json.loads(meta.read_text())
    t0 = time.time()
    per = 250_000
    jobs = [(i, min(per, reads - i), 1000 + i // per) for i in range(0, reads, per)]
    inflated = 0
    with open(path, "wb") as f, ProcessPoolExecutor(max_workers=min(32, os.cpu_count() or 8)) as ex:
        for blob, n in ex.map(bgzf_blocks, jobs):      # (a job ends its last member early: members are <= 0xff00 bytes, as BGZF requires)
            f.write(blob); inflated += n
        f.write(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
    info = {"reads": reads, "inflated_bytes": inflated, "compressed_bytes": path.stat().st_size, "generate_s": round(time.time() - t0, 1)}
    meta.write_text(json.dumps(info))
    return path, info


def main():
    import bamscan
    from oracle.fastq_oracle import parse_records
    path, info = ensure_file()
    print(f"[fastq] {path}: {info}", file=sys.stderr, flush=True)
    peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()).get("hbm_gbs", 6543.4) if (ROOT / "MEASURED_PEAKS.json").exists() else 6543.4
    prov = bamscan.FastqTableProvider(str(path))
    rows = []
    for tp in (1, 8):
        plan = prov.scan(None, None, None, target_partitions=tp)
        n = plan.output_partition_count()
        for rep in range(2):      # second repetition is reported (buffers warm)
            dev_ms, slow, nrows, stats = 0.0, 0.0, 0, []
            for p in range(n):
                st = plan.run_device_resident(p, 1)
                dev_ms += st["ms_total"]; slow = max(slow, st["ms_total"]); nrows += st["rows"]; stats.append(st)
            t0 = time.perf_counter(); e2e_rows = 0
            for p in range(n):
                for b in plan.execute(p):
                    e2e_rows += b.num_rows
                    del b
            e2e_ms = (time.perf_counter() - t0) * 1e3
        assert nrows == reads == e2e_rows, (nrows, e2e_rows, reads)
        if n > 1:
            bamscan.check_partition_seams(stats)
        infl_ms = sum(s["ms_inflate"] for s in stats)
        alg = sum(s["inflated_bytes"] + s["compressed_bytes"] for s in stats)
        rows.append((tp, nrows, dev_ms, slow, e2e_ms, infl_ms, alg, sum(s["arrow_bytes"] for s in stats)))
        print(f"[fastq] partitions={n} rows={nrows} device {dev_ms:.1f} ms (slowest partition {slow:.1f}) e2e {e2e_ms:.1f} ms", file=sys.stderr, flush=True)
    # CPU restatement on a sample (one thread: zlib inflate of the sample's members + the python record parser)
    sample = 500_000
    from oracle.fastq_oracle import gunzip_members
    raw = open(path, "rb").read(int(info["compressed_bytes"] * sample / reads * 1.02))
    t0 = time.perf_counter()
    txt, off = [], 0
    while off + 18 <= len(raw):                      # member by member through the BSIZE field (no quadratic tail copies)
        size = struct.unpack_from("<H", raw, off + 16)[0] + 1
        if off + size > len(raw):
            break
        txt.append(zlib.decompress(raw[off + 18:off + size - 8], -15))
        off += size
    text = b"".join(txt)
    text = text[:text.rfind(b"\n@SIM") + 1]
    recs = parse_records(text)
    cpu_s = time.perf_counter() - t0
    print("| partitions | rows | device ms (sum) | slowest partition ms | reads/s (device) | inflated GB/s (device) | inflate launches: GB/s of (csize + isize) / fraction of HBM peak | e2e ms | reads/s (e2e) |")
    print("|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    for tp, nrows, dev, slow, e2e, infl_ms, alg, arrow in rows:
        g = alg / (infl_ms / 1e3) / 1e9
        print(f"| {tp} | {nrows} | {dev:.1f} | {slow:.1f} | {nrows / dev * 1e3 / 1e6:.1f} M | {info['inflated_bytes'] / dev * 1e3 / 1e9:.1f} | {g:.1f} / {g / peak:.4f} | {e2e:.1f} | {nrows / e2e * 1e3 / 1e6:.1f} M |")
    print(f"\nfile: {reads} reads x {READ_LEN} bp, {info['inflated_bytes'] / 1e9:.2f} GB of text, {info['compressed_bytes'] / 1e9:.2f} GB BGZF (zlib level 6); Arrow bytes {rows[0][7] / 1e9:.2f} GB.")
    print(f"CPU restatement (oracle/fastq_oracle.py: zlib + python parser, 1 thread) on the first {len(recs)} reads: {len(recs) / cpu_s / 1e6:.2f} M reads/s.")
    print("Reference's own adjacent-path number (SURVEY 5, other hardware, 26.5 M reads): 1.5 M reads/s on one core, ~11 M reads/s on eight.")


if __name__ == "__main__":
    main()
