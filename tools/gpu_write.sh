#!/bin/bash
# write path on the GPU box: parity tests, plain measurement, then (only after the plain run exited 0) the ncu launch list and one
# --set full capture each of the two dominant kernels
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_write.py -x -q -m gpu 2>&1 | tail -3
BAMSCAN_WRITER_TRACE=1 timeout 300 python tools/measure_write.py 4000000 > gpurun_out/write_4m.json 2> gpurun_out/write_4m.err || exit 1
tail -12 gpurun_out/write_4m.err; cat gpurun_out/write_4m.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/write_launches.csv python tools/measure_write.py 1000000 > gpurun_out/write_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bgzf_deflate_kernel -s 1 -c 1 -f -o gpurun_out/w_deflate python tools/measure_write.py 2000000 >> gpurun_out/write_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:enc_records_kernel -s 1 -c 1 -f -o gpurun_out/w_encode python tools/measure_write.py 2000000 >> gpurun_out/write_ncu.log 2>&1
ls -la gpurun_out/w_*.ncu-rep
