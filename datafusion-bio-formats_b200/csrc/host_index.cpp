// host_index.cpp -- indexed (BAI) planning.  Filled in by the indexed-scan milestone; until then the planner
// falls back to the sequential single-partition scan (which returns the same row set for full scans).
#include "bamscan_internal.h"

namespace bamscan {

int plan_indexed(BamFile*, Plan*, const BamScanFilter*, int32_t, int32_t, bool* handled) {
  *handled = false;
  return BAMSCAN_OK;
}

}  // namespace bamscan
