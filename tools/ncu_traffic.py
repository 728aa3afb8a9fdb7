#!/usr/bin/env python3
"""DRAM traffic of the inflate kernels from one `ncu --set full` capture each, written to profiles/r2_inflate_traffic.json
together with the hash of the kernel sources it was taken on (bench.py reports `roofline.traffic` from this file and only
while the hash still matches the tree -- never a literal).

Run on the GPU box AFTER the plain command has exited 0:   python tools/prof_inflate.py 4500000 && python tools/ncu_traffic.py 4500000
(4.5 M reads = 23 650 members; bench.py scales the per-launch figure by members when its launches differ).
"""
import csv, hashlib, io, json, os, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
reads = sys.argv[1] if len(sys.argv) > 1 else "4500000"
SRC = ["kernels_inflate.cuh", "kernels_inflate_cta.cuh", "inflate_cta_core.h"]
sha = hashlib.sha256(b"".join((ROOT / "datafusion-bio-formats_b200" / "csrc" / k).read_bytes() for k in SRC)).hexdigest()[:16]
out = {"source_sha": sha, "reads": int(reads), "how": "ncu --set full --clock-control none, second launch of bamscan_bench_inflate (one full wave)", "kernels": {}}
(ROOT / "gpurun_out").mkdir(exist_ok=True)
for label, flags, regex in (("inflate_lg_kernel", "8", "regex:inflate_lg"), ("inflate_cta_kernel", "16", "regex:inflate_cta")):
    rep = ROOT / "gpurun_out" / f"traffic_{label}"
    env = dict(os.environ, BAMSCAN_DEBUG_FLAGS=flags)
    subprocess.run(["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "-k", regex, "-s", "1", "-c", "1", "-f", "-o", str(rep),
                    sys.executable, str(ROOT / "tools" / "prof_inflate.py"), reads], env=env, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    raw = subprocess.run(["ncu", "-i", str(rep) + ".ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    get = lambda name: (float(vals[hdr.index(name)]), units[hdr.index(name)])
    def to_bytes(name):
        v, u = get(name)
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
    dur, du = get("gpu__time_duration.sum")
    out["kernels"][label] = {"dram_read_bytes": rd, "dram_write_bytes": wr, "dram_bytes_per_launch": rd + wr,
                             "duration_ms_under_ncu": dur * {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}[du],
                             "registers": get("launch__registers_per_thread")[0], "ipc": get("sm__inst_executed.avg.per_cycle_elapsed")[0],
                             "warps_active_pct": get("sm__warps_active.avg.pct_of_peak_sustained_active")[0]}
# the product path uses the CTA-per-member kernel for every launch (the lane-group kernel is captured beside it as the A/B baseline)
out["dram_bytes_per_launch"] = out["kernels"]["inflate_cta_kernel"]["dram_bytes_per_launch"]
out["kernel"] = "inflate_cta_kernel (one launch over %d reads' members)" % int(reads)
(ROOT / "profiles").mkdir(exist_ok=True)
(ROOT / "profiles" / "r2_inflate_traffic.json").write_text(json.dumps(out, indent=1))
(ROOT / "gpurun_out" / "r2_inflate_traffic.json").write_text(json.dumps(out, indent=1))
print(json.dumps(out))
