#!/bin/bash
mkdir -p gpurun_out
bash tests/run_gpu_tests.sh > /dev/null; tail -2 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; tail -1 gpurun_out/smoke_final.log | cut -c1-200
python tools/measure_write.py 4000000 > gpurun_out/write_4m_final.json 2>/dev/null && cat gpurun_out/write_4m_final.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bgzf_deflate_kernel -s 1 -c 1 -f -o gpurun_out/w_deflate_v2 python tools/measure_write.py 2000000 > gpurun_out/write_ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:enc_records_kernel -s 1 -c 1 -f -o gpurun_out/w_encode_v2 python tools/measure_write.py 2000000 >> gpurun_out/write_ncu2.log 2>&1
ls -la gpurun_out/w_*_v2.ncu-rep
