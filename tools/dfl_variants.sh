#!/bin/bash
# A/B of compile-time variants of bgzf_deflate_kernel (builds on the GPU box): deflate stage of a 4 M-read write
cd datafusion-bio-formats_b200/csrc
for v in "$@"; do
  echo "== $v"
  rm -f ../libbamscan.so
  make CXXFLAGS="-O3 -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function -lineinfo $v" > /dev/null 2>&1 || { echo build failed; tail -5 build.log; continue; }
  grep -A1 "bgzf_deflate_kernel" build.log | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores" | head -2 | tr '\n' ' '; echo
  (cd ../.. && timeout 300 python tools/measure_write.py 4000000 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('deflate_ms', round(d['deflate_ms'],2), 'encode_ms', round(d['encode_ms'],2), 'size_vs_zlib6', round(d['size_vs_zlib6'],4), 'e2e', round(d['e2e_reads_per_s']/1e6,1), 'read_back', d['read_back_rows'])")
done
rm -f ../libbamscan.so; make > /dev/null 2>&1
