#!/bin/bash
# host-side variants of the writer (builds on the GPU box): end-to-end rate of a 4 M-read write, best of the tool's 3 repetitions is its last
cd datafusion-bio-formats_b200/csrc
for v in "$@"; do
  echo "== $v"
  rm -f ../libbamscan.so
  make CXXFLAGS="-O3 -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function -lineinfo $v" > /dev/null 2>&1 || { echo build failed; continue; }
  (cd ../.. && for k in 1 2; do timeout 300 python tools/measure_write.py 4000000 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('e2e', round(d['e2e_reads_per_s']/1e6,2), 'M reads/s', round(d['e2e_s']*1e3,1), 'ms')"; done)
done
rm -f ../libbamscan.so; make > /dev/null 2>&1
