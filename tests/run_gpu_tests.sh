#!/bin/bash
# usage (under gpurun): bash tests/run_gpu_tests.sh [pytest args]
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q "$@" 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
rc=${PIPESTATUS[0]}
tail -40 gpurun_out/pytest_gpu.log
exit $rc
