// cpp_mirror_check.cpp -- compile-and-run check of include/bamscan.hpp (planning calls only: no GPU needed).
//   cpp_mirror_check <multi_chrom.bam> <10x_pbmc_tags.bam> <sample.fastq.bgz (with its .gzi)>
#include <cstdio>
#include <cstdlib>

#include "bamscan.hpp"

using namespace bamscan_cpp;

#define REQUIRE(c) do { if (!(c)) { fprintf(stderr, "FAILED line %d: %s\n", __LINE__, #c); return 1; } } while (0)

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  BamTableProvider t(argv[1], std::nullopt, true, std::vector<std::string>{"NM", "MD"});
  ArrowSchema sc;
  t.schema(&sc);
  REQUIRE(sc.n_children == 14 && std::string(sc.children[0]->name) == "name" && std::string(sc.children[12]->name) == "NM");
  sc.release(&sc);
  std::vector<Filter> fs = {Filter::str(BAMSCAN_COL_CHROM, BAMSCAN_OP_EQ, {"chr1"}), Filter::num(BAMSCAN_COL_START, BAMSCAN_OP_BETWEEN, {100, 100000}),
                            Filter::num(0, BAMSCAN_OP_OTHER, {})};
  std::vector<bool> push = t.supports_filters_pushdown(fs);
  REQUIRE(push.size() == 3 && push[0] && push[1] && !push[2]);
  BamExec plan = t.scan(std::vector<int32_t>{1, 2, 3}, fs, std::nullopt, 4);
  REQUIRE(plan.output_partition_count() >= 1 && plan.output_partition_count() <= 4);
  plan.schema(&sc);
  REQUIRE(sc.n_children == 3 && std::string(sc.children[0]->name) == "chrom");
  sc.release(&sc);
  // unsatisfiable conjunction => EmptyExec (table_provider.rs:1005-1010)
  std::vector<Filter> none = {Filter::str(BAMSCAN_COL_CHROM, BAMSCAN_OP_EQ, {"chr1"}), Filter::num(BAMSCAN_COL_START, BAMSCAN_OP_GT, {1000}),
                              Filter::num(BAMSCAN_COL_START, BAMSCAN_OP_LT, {500})};
  REQUIRE(t.scan(std::nullopt, none, std::nullopt, 4).output_partition_count() == 0);
  REQUIRE(t.scan(std::nullopt, {}, std::nullopt, 8, BAMSCAN_PARTITION_BLOCK_RANGE).output_partition_count() == 8);
  BamTableProvider x = BamTableProvider::try_new_with_inferred_schema(argv[2], std::nullopt, true, std::vector<std::string>{"CB", "xf"}, std::nullopt, false);
  auto rows = x.describe_tags(50);
  REQUIRE(rows.size() == 14 && rows[0].tag == "AS" && rows[0].sam_type == "i" && rows[0].arrow_type == "Int32" && rows[0].description == "Alignment score");
  REQUIRE(rows.back().tag == "xf" && rows.back().description == "Custom/unknown tag (i)");
  bool threw = false;
  try { BamTableProvider bad("/nonexistent/file.bam"); } catch (const Error& e) { threw = e.code == BAMSCAN_ERR_IO; }
  REQUIRE(threw);
  try { BamTableProvider remote(argv[1], std::string("{}")); threw = false; } catch (const Error& e) { threw = e.code == BAMSCAN_ERR_UNSUPPORTED; }
  REQUIRE(threw);
  FastqTableProvider fq(argv[3]);
  fq.schema(&sc);
  REQUIRE(sc.n_children == 4 && std::string(sc.children[1]->name) == "description");
  sc.release(&sc);
  REQUIRE(fq.scan(std::nullopt, std::nullopt, 1).output_partition_count() == 1);
  REQUIRE(fq.scan(std::nullopt, std::nullopt, 4).output_partition_count() == 4);          // the reference's cut by block count from the .gzi
  REQUIRE(fq.scan(std::vector<int32_t>{0}, std::nullopt, 64).output_partition_count() == 10);   // min(target, blocks): 9 GZI entries + the first member
  if (argc > 100) {   // compile-only here: execute / write need a GPU (tests/test_gpu_*.py drive them through the same C ABI)
    RecordBatchStream st = plan.execute(0);
    ArrowArray batch;
    t.schema(&sc);
    BamWriteExec w("/tmp/out.bam", "@HD\tVN:1.6\n", {"chr1"}, {1000}, &sc, {"NM"});
    while (st.next(&batch)) { w.write(&batch); batch.release(&batch); }
    printf("%llu rows, %llu members\n", (unsigned long long)w.finish(), (unsigned long long)w.stats().members);
  }
  printf("cpp mirror ok\n");
  return 0;
}
