#!/bin/bash
# A/B of the two inflate kernels on one full wave of the cached bench file: 4 = warp per member, 8 = lane group.
reads=${1:-12000000}
for v in 4 8; do
  echo "== debug_flags $v"
  BAMSCAN_DEBUG_FLAGS=$v timeout 300 python tools/prof_inflate.py $reads full 2>&1 | tail -2
done
