#!/bin/bash
# Quick check of a changed CTA inflate kernel: parity tests forced onto it, then kernel-only timing with and without the phase profile.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cta_per_member" > gpurun_out/parity_cta.log 2>&1; echo "parity rc=$?" | tee -a gpurun_out/parity_cta.log
tail -4 gpurun_out/parity_cta.log
BAMSCAN_DEBUG_FLAGS=16 timeout 600 python tools/prof_inflate.py ${1:-4000000} 2>&1 | tail -1 | tee gpurun_out/icta_time.log
BAMSCAN_DEBUG_FLAGS=16 BAMSCAN_ICTA_PROF=1 timeout 600 python tools/prof_inflate.py ${1:-4000000} 2>&1 | tail -2 | tee -a gpurun_out/icta_time.log
