// host_planner_driver.cpp -- test-only driver of the product's HOST code (host_file / host_schema / host_index / host_plan .cpp:
// BGZF walk, BAM header, tag inference, BAI / CSI / GZI parsers, the planner), linked WITHOUT the CUDA engine so that it can be
// built with -fsanitize=address,undefined.  The few helpers engine.cu provides to those files are stubbed with malloc.
//   host_planner_driver file [index | "" | -] [fastq]
// Exit code 0 whether the input is accepted or refused; a sanitizer report (or a crash) is the failure the test looks for.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>

#include "../../datafusion-bio-formats_b200/csrc/bamscan_internal.h"

namespace bamscan {
void* pinned_alloc(size_t bytes) { return calloc(1, bytes); }
void pinned_free(void* p) { free(p); }
void* map_file_pinned(const char*, uint64_t, bool*) { return nullptr; }   // never taken: load_file falls back to a private copy
void unmap_file_pinned(void*, uint64_t, bool) {}
}  // namespace bamscan

using namespace bamscan;

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  const bool fastq = argc > 3 && !strcmp(argv[3], "fastq");
  BamFile f;
  f.path = argv[1]; f.format = fastq ? 1 : 0;
  f.has_tag_fields = true; f.tag_fields = {"NM", "MD", "xf", "pa", "ZZ", "CB"};
  int rc = load_file(&f);
  if (f.size >= (1u << 20)) { fprintf(stderr, "driver: files of 1 MiB and more are mapped by the engine; use a smaller input\n"); return 2; }
  if (rc == BAMSCAN_ERR_IO || rc == BAMSCAN_ERR_CUDA) { release_file(&f); return 0; }
  int accepted = rc == BAMSCAN_OK;
  if (!fastq) {
    BamScanOptions opt; memset(&opt, 0, sizeof opt);
    opt.struct_size = sizeof opt; opt.coordinate_system_zero_based = 1; opt.infer_tag_types = 1; opt.infer_tag_sample_size = 100; opt.has_tag_fields = 1;
    if (build_schema(&f, &opt) != BAMSCAN_OK) { release_file(&f); return 0; }
    if (f.header_ok) { std::map<std::string, std::pair<char, int32_t>> all; infer_tag_types(f, {}, 1000, &all, true); }
    { struct ArrowSchema sc; memset(&sc, 0, sizeof sc); if (export_schema(f.fields, f.schema_metadata, &sc) == BAMSCAN_OK && sc.release) sc.release(&sc); }   // Arrow C Data Interface: export + release (LeakSanitizer watches)
    if (argc > 2 && strcmp(argv[2], "-") != 0) {
      f.index_path = *argv[2] ? std::string(argv[2]) : discover_index(f.path);
      if (!f.index_path.empty()) {
        f.bai.reset(new BaiIndex());
        if (load_bai(f.index_path, f.bai.get()) != BAMSCAN_OK) { f.bai.reset(); f.index_path.clear(); }
      }
    }
  }
  if (f.header_ok) {
    const char* chroms[] = {"chr1", "chr2", "chrX", "nope"};
    for (int q = 0; q < 6; q++) {
      const char* cs[1] = {chroms[q % 4]};
      double between[2] = {1000.0 * q, 50000.0 * (q + 1)};
      BamScanFilter fl[2];
      fl[0] = BamScanFilter{BAMSCAN_COL_CHROM, 0 /* EQ */, 1, nullptr, cs};
      fl[1] = BamScanFilter{BAMSCAN_COL_START, 6 /* BETWEEN */, 2, between, nullptr};
      for (int tp : {1, 3, 16}) for (int mode : {BAMSCAN_PARTITION_REFERENCE, BAMSCAN_PARTITION_BLOCK_RANGE}) {
        Plan* p = nullptr;
        if (make_plan(&f, nullptr, -1, q ? fl : nullptr, q == 0 ? 0 : (q < 3 ? 1 : 2), tp, mode, &p) == BAMSCAN_OK && p) { accepted += (int)p->partitions.size(); delete p; }
      }
    }
  }
  release_file(&f);
  printf("driver: %s (%d)\n", accepted ? "accepted" : "refused", accepted);
  return 0;
}
