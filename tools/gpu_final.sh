#!/bin/bash
# round-end evidence on one B200: full parity suite, smoke, the default bench line (100 M reads, verify + write_path keys),
# the per-config table at 100 M reads
mkdir -p gpurun_out
bash tests/run_gpu_tests.sh > /dev/null; tail -2 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; tail -3 gpurun_out/smoke_final.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_100m_final.json 2> gpurun_out/bench_100m_final.err; tail -2 gpurun_out/bench_100m_final.err
python tools/measure_configs.py 100000000 200000 > gpurun_out/configs_final.md 2> gpurun_out/configs_final.err
cat gpurun_out/configs_final.md
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_100m_final.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"], "stage", d["stage_ms_rank0"])
print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "traffic", "ms_per_launch")}, d["roofline"]["kernel_alone"])
print("verify", d.get("verify")); print("write", d.get("write_path")); print("cpu", d.get("cpu_baseline")); print("clocks", d.get("clocks"))
PY
