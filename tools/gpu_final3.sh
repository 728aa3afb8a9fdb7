#!/bin/bash
mkdir -p gpurun_out
bash tests/run_gpu_tests.sh > /dev/null; tail -2 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; tail -1 gpurun_out/smoke_final.log | cut -c1-160
BAMSCAN_BENCH_READS=10000000 python bench.py --steps 3 --warmup 3 > gpurun_out/bench10m_final.json 2> gpurun_out/bench10m_final.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench10m_final.json").read().strip().splitlines()[-1])
print("value", round(d["value"] / 1e6, 1), "e2e", round(d["e2e"]["value"] / 1e6, 1), "frac", round(d["roofline"]["frac"], 4), "traffic", d["roofline"]["traffic"])
print("verify", d["verify"]["result"], d["verify"]["seconds"])
w = d["write_path"]; print("write", {k: w[k] for k in ("rows", "e2e_reads_per_s", "device_reads_per_s", "compression_ratio", "read_back")})
PY
