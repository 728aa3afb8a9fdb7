#!/bin/bash
# First-contact run for a changed inflate kernel: smoke, the fixture parity tests of the default kernel, then a kernel-only timing.
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cta_per_member" > gpurun_out/parity_cta.log 2>&1; echo "parity rc=$?" | tee -a gpurun_out/parity_cta.log
for v in 0 8; do
  echo "== debug_flags $v" >> gpurun_out/inflate_ab.log
  BAMSCAN_DEBUG_FLAGS=$v timeout 600 python tools/prof_inflate.py ${1:-4000000} full >> gpurun_out/inflate_ab.log 2>&1
done
tail -5 gpurun_out/smoke.log; tail -15 gpurun_out/parity_cta.log; cat gpurun_out/inflate_ab.log | tail -12
