#!/bin/bash
# A/B of the inflate kernel variants (BAMSCAN_INFLATE_VARIANT) on the cached bench file.
reads=${1:-8000000}
for v in ${VARIANTS:-0 1 2 3 4 5 6 7 8}; do
  echo "== variant $v"
  BAMSCAN_INFLATE_VARIANT=$v timeout 300 python tools/prof_inflate.py $reads full 2>&1 | tail -2
done
