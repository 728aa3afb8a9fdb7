#!/bin/bash
# configs 3 / 4 / 5 across 4 physical GPUs (strong scaling of one file: block-range partitions, or balance_partitions regions for config 4)
mkdir -p gpurun_out
export BAMSCAN_BENCH_READS=20000000
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 3 --warmup 2 "$@"; }
python bench.py --config 4 --steps 3 --warmup 2 > gpurun_out/n1_cfg4.json 2> gpurun_out/n1_cfg4.err
run --config 4 > gpurun_out/n4_cfg4.json 2> gpurun_out/n4_cfg4.err
run --config 3 --no-replicas > gpurun_out/n4_cfg3.json 2> gpurun_out/n4_cfg3.err
python bench.py --config 3 --steps 3 --warmup 2 > gpurun_out/n1_cfg3.json 2> gpurun_out/n1_cfg3.err
BAMSCAN_BENCH_LONG_READS=200000 run --config 5 > gpurun_out/n4_cfg5.json 2> gpurun_out/n4_cfg5.err
python - <<'PY'
import json
for n in ("n1_cfg4", "n4_cfg4", "n1_cfg3", "n4_cfg3", "n4_cfg5"):
    try:
        d = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, "value", round(d["value"]), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]), "rows", d["config"]["sample_reads_per_step"], [round(r["ms_per_step"], 2) for r in d["per_rank"]])
    except Exception as e:
        print(n, "FAILED", e)
PY
