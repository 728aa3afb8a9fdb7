#!/bin/bash
mkdir -p gpurun_out
bash tests/run_gpu_tests.sh
for f in 0 64; do
  BAMSCAN_DEBUG_FLAGS=$f python bench.py --config 5 --reads 200000 --steps 3 --warmup 2 > gpurun_out/cfg5_200k_f$f.json 2> gpurun_out/cfg5_200k_f$f.err
  python - <<PY
import json
d=json.load(open("gpurun_out/cfg5_200k_f$f.json"))
print("flags $f:", d["value"], d["stage_ms_rank0"], "repairs", d["per_rank"][0]["boundary_repairs"], "mismatches", d["per_rank"][0]["seam_mismatches"], "launches", d["gpu_launches"])
PY
done
