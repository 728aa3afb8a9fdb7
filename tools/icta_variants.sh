#!/bin/bash
# A/B of compile-time variants of the CTA-per-member inflate kernel on one wave (builds on the GPU box: nvcc is there).
# usage: tools/icta_variants.sh reads "flags1" "flags2" ...
reads=${1:-4000000}; shift
cd datafusion-bio-formats_b200/csrc
for v in "$@"; do
  echo "== $v"
  rm -f ../libbamscan.so
  make CXXFLAGS="-O3 -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function -lineinfo $v" > /dev/null 2>&1 || { echo build failed; continue; }
  (cd ../.. && BAMSCAN_DEBUG_FLAGS=16 BAMSCAN_ICTA_PROF=1 timeout 300 python tools/prof_inflate.py $reads 2>&1 | tail -2)
done
rm -f ../libbamscan.so; make > /dev/null 2>&1
