"""Whole-file parity check of a GPU scan against the CPU oracle (test infrastructure: imported by tests/ and by
`bench.py --verify` only, never by the product).

The oracle scans the file as block-range partitions cut at exact record starts (BAI linear-index offsets), one host thread
per partition; the GPU scan arrives as whatever batches the engine produces.  Because the two sides chunk the rows
differently, equality is established in two ways:
  * per column, a streaming 64-bit checksum (CRC-32 and Adler-32 chained in row order) of the LOGICAL content -- validity
    as one byte per row, fixed-width values with nulls zeroed, var-len lengths, and the concatenated value bytes -- over
    the whole file on both sides;
  * full RecordBatch equality on the first, a middle and the last oracle partition against the same row ranges of the
    GPU output.
Reference semantics being checked: physical_exec.rs:371-598 (sequential scan), alignment_utils.rs:316-368 (batch assembly).
"""
import struct
import zlib

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc


def bai_linear_offsets(bai_path) -> list:
    """Sorted distinct linear-index virtual offsets of a .bai: every one is the start of a record."""
    data = open(bai_path, "rb").read()
    assert data[:4] == b"BAI\1"
    n_ref, = struct.unpack_from("<i", data, 4)
    pos, out = 8, set()
    for _ in range(n_ref):
        n_bin, = struct.unpack_from("<i", data, pos); pos += 4
        for _ in range(n_bin):
            _b, n_chunk = struct.unpack_from("<Ii", data, pos); pos += 8 + 16 * n_chunk
        n_intv, = struct.unpack_from("<i", data, pos); pos += 4
        out.update(struct.unpack_from(f"<{n_intv}Q", data, pos)); pos += 8 * n_intv
    out.discard(0)
    return sorted(out)


class ColumnSum:
    """Streaming (crc32, adler32) over the logical content of one column, chunking-independent."""

    def __init__(self):
        self.rows = 0
        self.nulls = 0
        self.state = {"validity": [0, 1], "values": [0, 1], "lengths": [0, 1]}

    def _feed(self, key, buf):
        s = self.state[key]
        s[0] = zlib.crc32(buf, s[0]); s[1] = zlib.adler32(buf, s[1])

    def update(self, arr: pa.Array):
        n = len(arr)
        self.rows += n; self.nulls += arr.null_count
        self._feed("validity", np.asarray(arr.is_valid().to_numpy(zero_copy_only=False), dtype=np.uint8).tobytes())
        t = arr.type
        if pa.types.is_string(t) or pa.types.is_binary(t):
            lens = pc.fill_null(pc.binary_length(arr), 0).to_numpy(zero_copy_only=False).astype(np.int32)
            self._feed("lengths", lens.tobytes())
            # value bytes of the non-null rows, in order (null slots have zero length in both producers)
            offs = np.frombuffer(arr.buffers()[1], dtype=np.int32, count=n + 1, offset=arr.offset * 4)
            data = arr.buffers()[2]
            if data is not None and n:
                self._feed("values", memoryview(data)[int(offs[0]):int(offs[-1])])
        elif pa.types.is_list(t):
            lens = pc.fill_null(pc.list_value_length(arr), 0).to_numpy(zero_copy_only=False).astype(np.int32)
            self._feed("lengths", lens.tobytes())
            self._feed("values", arr.flatten().to_numpy(zero_copy_only=False).tobytes())
        else:
            zero = pa.scalar(0, type=t) if not pa.types.is_floating(t) else pa.scalar(0.0, type=t)
            self._feed("values", pc.fill_null(arr, zero).to_numpy(zero_copy_only=False).tobytes())

    def digest(self):
        return (self.rows, self.nulls) + tuple((v[0] << 32) | v[1] for v in (self.state["validity"], self.state["lengths"], self.state["values"]))


def oracle_partitions(oracle, bai_path, n_parts, projection=None, threads=None):
    """Scans the whole file with the oracle as n_parts block-range partitions (exact record starts) on host threads and
    yields the RecordBatches in file order.  At most `threads` partitions are alive at a time (a 100 M-read file does not
    fit in host memory as Arrow twice over)."""
    import concurrent.futures as cf
    offs = bai_linear_offsets(bai_path)
    n_parts = max(1, min(n_parts, len(offs)))
    cuts = [offs[(len(offs) * k) // n_parts] for k in range(1, n_parts)]
    cuts = [oracle.first_record_voffset] + sorted(set(cuts)) + [0]
    n = len(cuts) - 1
    threads = max(1, min(threads or n, n))
    with cf.ThreadPoolExecutor(max_workers=threads) as ex:
        futs = {}
        nxt = 0
        for i in range(n):
            while nxt < n and nxt < i + threads:
                futs[nxt] = ex.submit(oracle.scan, projection, cuts[nxt], 0, 0, 0, None, None, cuts[nxt + 1])
                nxt += 1
            yield futs.pop(i).result()


def verify_full_scan(oracle, bai_path, gpu_batches, n_parts=8, projection=None, threads=None) -> dict:
    """gpu_batches: iterable of RecordBatch in file order.  Raises AssertionError on any difference."""
    # pass 1 (oracle): checksums of every column over every row; the first, a middle and the last partition are kept
    picks = sorted({0, n_parts // 2, n_parts - 1})
    names, want, keep, starts, rows_of = None, None, {}, [], []
    row0 = 0
    for k, b in enumerate(oracle_partitions(oracle, bai_path, n_parts, projection, threads)):
        if names is None:
            names = b.schema.names
            want = {nm: ColumnSum() for nm in names}
        starts.append(row0); rows_of.append(b.num_rows); row0 += b.num_rows
        for nm in names:
            want[nm].update(b.column(nm))
        if k in picks:
            keep[k] = b
    total_rows = row0
    picks = sorted(keep)
    if len(starts) - 1 not in keep:                     # fewer partitions than asked for: the last one is whatever came last
        picks = sorted(set(picks))
    windows = {k: (starts[k], starts[k] + rows_of[k]) for k in picks}
    # pass 2 (GPU): same checksums, plus the rows of the kept windows
    got = {nm: ColumnSum() for nm in names}
    kept = {k: [] for k in picks}
    r = 0
    for b in gpu_batches:
        assert b.schema.names == names, f"column order {b.schema.names} != {names}"
        for nm in names:
            got[nm].update(b.column(nm))
        for k, (lo, hi) in windows.items():
            a, z = max(lo, r), min(hi, r + b.num_rows)
            if a < z:
                kept[k].append(b.slice(a - r, z - a))
        r += b.num_rows
    assert r == total_rows, f"rows: gpu {r} != oracle {total_rows}"
    for nm in names:
        assert got[nm].digest() == want[nm].digest(), f"column {nm}: checksum (rows, nulls, validity, lengths, values) gpu {got[nm].digest()} != oracle {want[nm].digest()}"
    for k in picks:
        g = pa.Table.from_batches(kept[k], schema=kept[k][0].schema) if kept[k] else None
        w = pa.Table.from_batches([keep[k]])
        assert g is not None and g.num_rows == w.num_rows, f"window {k}: rows"
        for nm in names:
            assert g[nm].combine_chunks().equals(w[nm].combine_chunks()), f"window {k} (rows {windows[k]}): column {nm} differs"
    return {"rows": total_rows, "columns": len(names), "oracle_partitions": len(starts),
            "windows_compared": [list(windows[k]) for k in picks],
            "checksums": {nm: [hex(x) for x in got[nm].digest()[2:]] for nm in names}}
