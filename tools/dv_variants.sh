#!/bin/bash
# A/B of compile-time variants of decode_var_kernel: decode stage of a 10 M-read device-resident scan (builds on the GPU box)
cd datafusion-bio-formats_b200/csrc
for v in "$@"; do
  echo "== $v"
  rm -f ../libbamscan.so
  make CXXFLAGS="-O3 -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function -lineinfo $v" > /dev/null 2>&1 || { echo build failed; continue; }
  grep -A1 "decode_var_kernelILi8" build.log | grep -o "Used [0-9]* registers" | head -1
  (cd ../.. && BAMSCAN_BENCH_READS=10000000 timeout 300 python bench.py --steps 5 --warmup 3 --no-verify 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['stage_ms_rank0'], round(d['value']/1e6,1))")
done
rm -f ../libbamscan.so; make > /dev/null 2>&1
