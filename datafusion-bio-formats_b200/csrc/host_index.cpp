// host_index.cpp -- indexed (BAI) planning: SQL predicates -> genomic regions -> balanced partitions -> BGZF chunk ranges.
//
// Replaces (reference, datafusion/):
//   bio-format-core/src/genomic_filter.rs:51-100,151-329    extract_genomic_regions / collect_genomic_constraints
//   bio-format-core/src/genomic_filter.rs:107-117           build_full_scan_regions
//   bio-format-bam/src/storage.rs:336-436                   estimate_sizes_from_bai
//   bio-format-bam/src/storage.rs:442-450                   unplaced_unmapped_count_from_bai
//   bio-format-core/src/partition_balancer.rs:61-295        balance_partitions
//   bio-format-bam/src/table_provider.rs:1001-1093          the indexed branch of scan()
//   bio-format-bam/src/physical_exec.rs:844-855,1131-1267   region -> noodles query / unmapped-tail seek position
//   noodles-csi BinningIndex::query (reg2bins, linear-index minimum offset, chunk merge; SAMv1 5.1.1-5.3)
#include <algorithm>
#include <cmath>
#include <cstring>

#include "bamscan_internal.h"

namespace bamscan {

bool filter_is_record_pushable(const BamFile& f, const BamScanFilter& flt);

// ---------------------------------------------------------------------------------------------
// genomic_filter.rs:151-329 over the flat conjunction of single-column predicates the ABI carries
struct RegionAnalysis {
  std::vector<std::string> chroms;
  bool has_lower = false, has_upper = false, unsatisfiable = false;
  uint64_t lower = 0, upper = 0;
  std::vector<uint8_t> residual;   // per filter: 1 = not consumed as a genomic constraint
};

static bool as_u64(double v, uint64_t* out) {   // scalar_to_u64: non-negative integer literals only
  if (!(v >= 0) || v > 1.8e19 || std::floor(v) != v) return false;
  *out = (uint64_t)v;
  return true;
}

static RegionAnalysis analyze_regions(const BamScanFilter* filters, int n, bool zero_based) {
  RegionAnalysis A;
  A.residual.assign((size_t)n, 1);
  auto set_lower = [&](uint64_t v) { A.lower = A.has_lower ? std::max(A.lower, v) : v; A.has_lower = true; };
  auto set_upper = [&](uint64_t v) { A.upper = A.has_upper ? std::min(A.upper, v) : v; A.has_upper = true; };
  for (int i = 0; i < n; i++) {
    const BamScanFilter& f = filters[i];
    const bool one_num = f.num_values && f.n_values >= 1, one_str = f.str_values && f.n_values >= 1;
    uint64_t v = 0, w = 0;
    if (f.column == BAMSCAN_COL_CHROM) {
      if (f.op == BAMSCAN_OP_EQ && one_str) { A.chroms.push_back(f.str_values[0]); A.residual[i] = 0; }
      else if (f.op == BAMSCAN_OP_IN && one_str) { for (int k = 0; k < f.n_values; k++) A.chroms.push_back(f.str_values[k]); A.residual[i] = 0; }
    } else if (f.column == BAMSCAN_COL_START) {
      if (f.op == BAMSCAN_OP_BETWEEN && f.num_values && f.n_values >= 2 && as_u64(f.num_values[0], &v) && as_u64(f.num_values[1], &w)) {
        set_lower(v + (zero_based ? 1 : 0)); set_upper(w + (zero_based ? 1 : 0)); A.residual[i] = 0;
      } else if (one_num && f.op <= BAMSCAN_OP_GE && as_u64(f.num_values[0], &v)) {
        uint64_t v1 = v + (zero_based ? 1 : 0);
        A.residual[i] = 0;
        switch (f.op) {
          case BAMSCAN_OP_EQ: set_lower(v1); set_upper(v1); break;
          case BAMSCAN_OP_GT: set_lower(v1 + 1); break;
          case BAMSCAN_OP_GE: set_lower(v1); break;
          case BAMSCAN_OP_LT: set_upper(v1 ? v1 - 1 : 0); break;
          case BAMSCAN_OP_LE: set_upper(v1); break;
          default: A.residual[i] = 1;   // != stays residual
        }
      }
    } else if (f.column == BAMSCAN_COL_END) {
      // `end` is 1-based inclusive in both coordinate systems and is never shifted (genomic_filter.rs:238-263)
      if (one_num && as_u64(f.num_values[0], &v)) {
        A.residual[i] = 0;
        switch (f.op) {
          case BAMSCAN_OP_EQ: case BAMSCAN_OP_LE: set_upper(v); break;
          case BAMSCAN_OP_LT: set_upper(v ? v - 1 : 0); break;
          default: A.residual[i] = 1;
        }
      }
    }
  }
  std::sort(A.chroms.begin(), A.chroms.end());
  A.chroms.erase(std::unique(A.chroms.begin(), A.chroms.end()), A.chroms.end());
  A.unsatisfiable = A.has_lower && A.has_upper && A.lower > A.upper;
  return A;
}

// ---------------------------------------------------------------------------------------------
// partition_balancer.rs:15-41
struct SizeEstimate {
  GenomicRegion region;
  uint64_t estimated_bytes = 0;
  uint64_t contig_length = 0;    // 0 = None
  uint64_t unmapped_count = 0;
  std::vector<uint64_t> nonempty_bin_positions;
  uint64_t leaf_bin_span = 0;
  int32_t source_index = -1;
};
struct Assignment { std::vector<GenomicRegion> regions; std::vector<int32_t> source; uint64_t total_estimated_bytes = 0; };

// partition_balancer.rs:61-295 (linear scan with per-partition byte budgets)
static std::vector<Assignment> balance_partitions(const std::vector<SizeEstimate>& estimates, size_t target_partitions) {
  std::vector<Assignment> partitions;
  if (estimates.empty()) return partitions;
  const size_t target = std::max<size_t>(target_partitions, 1);
  uint64_t total_bytes = 0;
  for (auto& e : estimates) total_bytes += e.estimated_bytes;
  if (target == 1) {
    Assignment a;
    for (auto& e : estimates) { a.regions.push_back(e.region); a.source.push_back(e.source_index); }
    a.total_estimated_bytes = total_bytes;
    partitions.push_back(a);
    return partitions;
  }
  if (total_bytes == 0) {   // round robin
    size_t nb = std::min(target, estimates.size());
    partitions.resize(nb);
    for (size_t i = 0; i < estimates.size(); i++) { partitions[i % nb].regions.push_back(estimates[i].region); partitions[i % nb].source.push_back(estimates[i].source_index); }
    return partitions;
  }
  const size_t effective_target = (size_t)std::min<uint64_t>(target, total_bytes);
  const uint64_t base_budget = total_bytes / effective_target;
  const size_t extra = (size_t)(total_bytes % effective_target);
  auto budget_for = [&](size_t idx) { return idx < extra ? base_budget + 1 : base_budget; };
  partitions.emplace_back();
  uint64_t budget = budget_for(0);
  for (const SizeEstimate& est : estimates) {
    uint64_t remaining = est.estimated_bytes;
    uint64_t eff_start = 0, eff_end = 0;
    if (est.region.has_start && est.region.has_end && est.region.end >= est.region.start) { eff_start = est.region.start; eff_end = est.region.end; }
    else if (est.contig_length > 0) { eff_start = 1; eff_end = est.contig_length; }
    const bool can_split = eff_end > 0 && eff_end >= eff_start;
    uint64_t pos = eff_start;
    bool was_split = false;
    if (remaining == 0) {   // still assigned: to the partition with the fewest regions
      size_t min_idx = 0;
      for (size_t i = 1; i < partitions.size(); i++) if (partitions[i].regions.size() < partitions[min_idx].regions.size()) min_idx = i;
      partitions[min_idx].regions.push_back(est.region); partitions[min_idx].source.push_back(est.source_index);
      continue;
    }
    while (remaining > 0) {
      if (budget == 0 && partitions.size() < effective_target) { partitions.emplace_back(); budget = budget_for(partitions.size() - 1); }
      const bool is_last = partitions.size() >= effective_target;
      const uint64_t remaining_bp = (can_split && pos <= eff_end) ? eff_end - pos + 1 : 0;
      const bool splittable = remaining_bp > 1;
      if (remaining <= budget || is_last || !splittable) {
        GenomicRegion r = est.region;
        if (can_split && pos <= eff_end && was_split) { r = GenomicRegion(); r.chrom = est.region.chrom; r.has_start = true; r.start = pos; }
        Assignment& p = partitions.back();
        p.regions.push_back(r); p.source.push_back(est.source_index);
        p.total_estimated_bytes += remaining;
        budget = budget > remaining ? budget - remaining : 0;
        remaining = 0;
      } else {
        was_split = true;
        uint64_t sub_end;
        auto bp_proportional = [&]() {
          uint64_t bp = (uint64_t)((unsigned __int128)remaining_bp * budget / remaining);
          bp = std::min(std::max<uint64_t>(bp, 1), remaining_bp - 1);
          return pos + bp - 1;
        };
        if (!est.nonempty_bin_positions.empty() && est.leaf_bin_span > 0) {
          const auto& B = est.nonempty_bin_positions;
          size_t range_start = (size_t)(std::lower_bound(B.begin(), B.end(), pos) - B.begin());
          size_t range_end = (size_t)(std::upper_bound(B.begin(), B.end(), eff_end) - B.begin());
          size_t bins_in_range = range_end - range_start;
          if (bins_in_range > 1) {
            size_t take = (size_t)((unsigned __int128)bins_in_range * budget / remaining);
            take = std::min(std::max<size_t>(take, 1), bins_in_range - 1);
            uint64_t bin_start = B[range_start + take - 1];
            sub_end = std::min(bin_start + est.leaf_bin_span - 1, eff_end - 1);
          } else sub_end = bp_proportional();
        } else sub_end = bp_proportional();
        GenomicRegion r; r.chrom = est.region.chrom; r.has_start = true; r.start = pos; r.has_end = true; r.end = sub_end;
        Assignment& p = partitions.back();
        p.regions.push_back(r); p.source.push_back(est.source_index);
        p.total_estimated_bytes += budget;
        remaining -= budget;
        pos = sub_end + 1;
        if (partitions.size() < effective_target) { partitions.emplace_back(); budget = budget_for(partitions.size() - 1); }
        else budget = 0;
      }
    }
    if (est.unmapped_count > 0) {
      GenomicRegion tail; tail.chrom = est.region.chrom; tail.unmapped_tail = true;
      Assignment& p = partitions.back();
      p.regions.push_back(tail); p.source.push_back(est.source_index);
      p.total_estimated_bytes += 1;
      budget = budget > 1 ? budget - 1 : 0;
    }
  }
  partitions.erase(std::remove_if(partitions.begin(), partitions.end(), [](const Assignment& a) { return a.regions.empty(); }), partitions.end());
  return partitions;
}

// storage.rs:336-436
static std::vector<SizeEstimate> estimate_sizes_from_bai(const BamFile& f, const std::vector<GenomicRegion>& regions) {
  std::vector<SizeEstimate> out;
  const BaiIndex& idx = *f.bai;
  for (size_t i = 0; i < regions.size(); i++) {
    SizeEstimate e;
    e.region = regions[i]; e.source_index = (int32_t)i;
    e.leaf_bin_span = 1ull << idx.min_shift;   // 16384 for a BAI
    int ref = -1;
    for (size_t k = 0; k < f.meta_ref_names.size(); k++) if (f.meta_ref_names[k] == regions[i].chrom) ref = (int)k;   // last duplicate wins (HashMap collect)
    e.estimated_bytes = 1;
    if (ref >= 0 && (size_t)ref < idx.refs.size()) {
      const BaiRef& R = idx.refs[(size_t)ref];
      uint64_t mn = ~0ull, mx = 0;
      for (auto& kv : R.bins) for (auto& c : kv.second) { mn = std::min(mn, c.beg >> 16); mx = std::max(mx, c.end >> 16); }
      e.estimated_bytes = mx > mn ? mx - mn : 0;
      e.unmapped_count = R.has_meta ? R.n_unmapped : 0;
      const uint32_t leaf0 = idx.leaf_first(), leaf1 = idx.level_first(idx.depth + 1);   // 4681 .. 37448 for a BAI
      for (auto& kv : R.bins) if (kv.first >= leaf0 && kv.first < leaf1) e.nonempty_bin_positions.push_back((uint64_t)(kv.first - leaf0) * e.leaf_bin_span + 1);
      std::sort(e.nonempty_bin_positions.begin(), e.nonempty_bin_positions.end());
    }
    if (ref >= 0 && (size_t)ref < f.meta_ref_lens.size() && f.meta_ref_lens[(size_t)ref] > 0) e.contig_length = f.meta_ref_lens[(size_t)ref];
    out.push_back(e);
  }
  return out;
}

// ---------------------------------------------------------------------------------------------
// BAI query: bins overlapping [beg, end) (0-based half open), linear-index minimum offset, merged chunks
// Chunks of every indexed bin overlapping [beg, end) (0-based half open).  reg2bins (SAMv1 5.3; CSIv1) names the candidate
// bins level by level -- level l holds bins of 2^(min_shift + 3 (depth - l)) bases with ids from (8^l - 1) / 7 --; instead of
// listing all candidates (8^depth leaf bins for an open-ended region of a deep CSI) each level is one range walk over the
// bins the index actually holds.
static void overlapping_chunks(const BaiIndex& I, const BaiRef& R, uint64_t beg, uint64_t end, std::vector<BaiChunk>* chunks) {
  if (end > I.max_pos()) end = I.max_pos();
  if (beg >= end) return;
  --end;
  for (int l = 0; l <= I.depth; l++) {
    const int sh = I.min_shift + 3 * (I.depth - l);
    const uint32_t t = I.level_first(l);
    const uint32_t lo = t + (uint32_t)(beg >> sh), hi = t + (uint32_t)(end >> sh);
    for (auto it = R.bins.lower_bound(lo); it != R.bins.end() && it->first <= hi; ++it) chunks->insert(chunks->end(), it->second.begin(), it->second.end());
  }
}

static std::vector<BaiChunk> bai_query(const BaiIndex& I, const BaiRef& R, uint64_t start1, bool has_end, uint64_t end1) {
  const uint64_t beg0 = start1 ? start1 - 1 : 0, end0 = has_end ? end1 : I.max_pos();
  std::vector<BaiChunk> chunks;
  overlapping_chunks(I, R, beg0, end0, &chunks);
  uint64_t min_off = 0;
  if (I.csi) {
    // noodles-csi BinnedIndex::min_offset: the loffset of the leaf bin holding the start, else of its nearest indexed ancestor
    uint32_t bin = I.leaf_first() + (uint32_t)(std::min(beg0, I.max_pos() - 1) >> I.min_shift);
    for (;;) {
      auto it = R.loffset.find(bin);
      if (it != R.loffset.end()) { min_off = it->second; break; }
      if (bin == 0) break;
      bin = (bin - 1) >> 3;
    }
  } else {
    size_t win = (size_t)(beg0 >> 14);
    if (!R.intervals.empty()) min_off = win < R.intervals.size() ? R.intervals[win] : R.intervals.back();
  }
  // Upper bound (this build; noodles reads every chunk of the overlapping bins): a BAI indexes a coordinate-sorted file, so
  // every record that starts at or before the region end precedes the first record held by a LEAF bin (16 kb window,
  // records lying entirely inside it) of a later window.  Chunks of the coarse bins beyond that offset only hold records
  // that start after the region and would be dropped by the row rule (physical_exec.rs:1295-1314) anyway.
  uint64_t max_off = ~0ull;
  if (has_end) {
    const uint32_t leaf0 = I.leaf_first(), n_leaf = 1u << (3 * I.depth);
    const uint32_t last_leaf = leaf0 + (uint32_t)std::min<uint64_t>((end1 ? end1 - 1 : 0) >> I.min_shift, n_leaf - 1);
    for (auto it = R.bins.upper_bound(last_leaf); it != R.bins.end() && it->first < leaf0 + n_leaf; ++it)
      for (auto& c : it->second) max_off = std::min(max_off, c.beg);
  }
  std::vector<BaiChunk> kept;
  for (auto& c : chunks) if (c.end > min_off && c.beg < max_off) kept.push_back(BaiChunk{std::max(c.beg, min_off), std::min(c.end, max_off)});
  std::sort(kept.begin(), kept.end(), [](const BaiChunk& a, const BaiChunk& b) { return a.beg < b.beg; });
  std::vector<BaiChunk> merged;
  for (auto& c : kept) {
    if (!merged.empty() && c.beg <= merged.back().end) merged.back().end = std::max(merged.back().end, c.end);
    else merged.push_back(c);
  }
  return merged;
}

static int32_t block_of_coff(const BamFile& f, uint64_t coff) {
  size_t lo = 0, hi = f.blocks.size();
  while (lo < hi) { size_t mid = (lo + hi) / 2; if (f.blocks[mid].coff < coff) lo = mid + 1; else hi = mid; }
  return (lo < f.blocks.size() && f.blocks[lo].coff == coff) ? (int32_t)lo : -1;
}

// virtual offset -> (block, inflated offset).  A virtual offset at/after EOF maps to total_inflated.
static bool voffset_to_uoff(const BamFile& f, uint64_t v, uint32_t* block, uint64_t* uoff) {
  uint64_t coff = v >> 16;
  if (coff >= f.size) { *block = (uint32_t)f.blocks.size(); *uoff = f.total_inflated; return true; }
  int32_t b = block_of_coff(f, coff);
  if (b < 0) return false;
  *block = (uint32_t)b; *uoff = f.blocks[(size_t)b].uoff + (v & 0xffff);
  return true;
}

static int add_range(const BamFile& f, uint64_t beg_v, uint64_t end_v_or_0, const ScanRange& rule, Partition* part) {
  uint32_t b0, b1; uint64_t u0, u1;
  if (!voffset_to_uoff(f, beg_v, &b0, &u0)) { set_error("index virtual offset %llu does not point at a BGZF block", (unsigned long long)beg_v); return BAMSCAN_ERR_FORMAT; }
  ScanRange r = rule;
  r.exact_start = true; r.first_uoff = u0; r.block_begin = b0;
  if (end_v_or_0) {
    if (!voffset_to_uoff(f, end_v_or_0, &b1, &u1)) { set_error("index virtual offset %llu does not point at a BGZF block", (unsigned long long)end_v_or_0); return BAMSCAN_ERR_FORMAT; }
    r.stop_uoff = u1;
    r.block_end = (b1 < f.blocks.size() && u1 > f.blocks[b1].uoff) ? b1 + 1 : b1;
    if (r.block_end <= r.block_begin) r.block_end = std::min<uint32_t>((uint32_t)f.blocks.size(), r.block_begin + 1);
    // one look-ahead member: the last owned record usually ends in the next block; inflating it in the same launch costs
    // nothing, a separate extension chunk costs a full single-member latency (records past stop_uoff are not owned)
    r.block_end = std::min<uint32_t>((uint32_t)f.blocks.size(), r.block_end + 1);
  } else { r.stop_uoff = ~0ull; r.block_end = (uint32_t)f.blocks.size(); }
  if (u0 >= f.total_inflated || r.block_begin >= f.blocks.size()) return BAMSCAN_OK;   // nothing to read
  if (r.stop_uoff != ~0ull && r.stop_uoff <= r.first_uoff) return BAMSCAN_OK;
  part->ranges.push_back(r);
  return BAMSCAN_OK;
}

int plan_indexed(BamFile* f, Plan* plan, const BamScanFilter* filters, int32_t n_filters, int32_t target_partitions, bool* handled) {
  *handled = false;
  if (!f->bai) return BAMSCAN_OK;
  RegionAnalysis A = analyze_regions(filters, n_filters, f->zero_based);
  if (A.unsatisfiable) { plan->empty_exec = true; *handled = true; return BAMSCAN_OK; }   // EmptyExec (table_provider.rs:1005-1010)
  std::vector<GenomicRegion> regions;
  const bool is_full_scan = A.chroms.empty();
  if (!A.chroms.empty()) {
    for (auto& c : A.chroms) { GenomicRegion r; r.chrom = c; r.has_start = A.has_lower; r.start = A.lower; r.has_end = A.has_upper; r.end = A.upper; regions.push_back(r); }
  } else {
    for (auto& n : f->meta_ref_names) { GenomicRegion r; r.chrom = n; regions.push_back(r); }   // build_full_scan_regions
  }
  if (regions.empty()) return BAMSCAN_OK;   // falls back to the sequential scan
  std::vector<SizeEstimate> est = estimate_sizes_from_bai(*f, regions);
  std::vector<Assignment> assignments = balance_partitions(est, (size_t)std::max(1, target_partitions));
  if (is_full_scan && f->bai->has_no_coor && f->bai->n_no_coor > 0) {   // table_provider.rs:285-306, 1048-1056
    Assignment a; GenomicRegion r; r.chrom = "*"; r.unmapped_tail = true;
    a.regions.push_back(r); a.source.push_back(-1); a.total_estimated_bytes = std::max<uint64_t>(f->bai->n_no_coor, 1);
    assignments.push_back(a);
  }
  for (int i = 0; i < n_filters; i++) {   // residual = every record-level pushable filter (table_provider.rs:1061-1065)
    if (!filter_is_record_pushable(*f, filters[i])) continue;
    RecordFilter rf; rf.column = filters[i].column; rf.op = filters[i].op;
    for (int k = 0; k < filters[i].n_values; k++) {
      if (filters[i].num_values) rf.nums.push_back(filters[i].num_values[k]);
      if (filters[i].str_values) rf.strs.push_back(filters[i].str_values[k]);
    }
    plan->residual.push_back(rf);
  }
  // An upper bound on `start` that the pushed record filters enforce on every row (start BETWEEN / = / < / <=).  balance_partitions
  // leaves the last sub-region of a split contig open-ended (partition_balancer.rs: `end = None`), so the reference -- and this
  // build before -- reads that partition's chunks up to the end of the contig and throws the rows away again in the filter.
  // The file is coordinate sorted: chunks that only hold records starting behind the bound cannot contribute a row, so they are
  // not read (same argument as the upper-bound prune of bai_query; fewer bytes, identical rows).
  bool has_start_sup = false; uint64_t start_sup = 0;
  for (int i = 0; i < n_filters; i++) {
    const BamScanFilter& q = filters[i];
    if (q.column != BAMSCAN_COL_START || !q.num_values || !filter_is_record_pushable(*f, q)) continue;
    uint64_t v = 0; bool have = false;
    if (q.op == BAMSCAN_OP_BETWEEN && q.n_values >= 2) have = as_u64(q.num_values[1], &v);
    else if ((q.op == BAMSCAN_OP_EQ || q.op == BAMSCAN_OP_LE) && q.n_values >= 1) have = as_u64(q.num_values[0], &v);
    else if (q.op == BAMSCAN_OP_LT && q.n_values >= 1) { have = as_u64(q.num_values[0], &v); if (have) v = v ? v - 1 : 0; }
    if (!have) continue;
    v += f->zero_based ? 1 : 0;                       // 1-based closed, like GenomicRegion
    start_sup = has_start_sup ? std::min(start_sup, v) : v; has_start_sup = true;
  }
  for (const Assignment& a : assignments) {
    Partition part;
    part.regions = a.regions; part.estimated_bytes = a.total_estimated_bytes;
    for (const GenomicRegion& g : a.regions) {
      ScanRange rule;
      if (g.unmapped_tail && g.chrom == "*") {           // physical_exec.rs:1038-1129: full re-scan keeping refID == -1 && pos == -1
        rule.region_mode = 3; rule.region_ref = -1;
        uint32_t b; uint64_t u;
        (void)b; (void)u;
        ScanRange r = rule; r.exact_start = true; r.first_uoff = f->first_record_uoff; r.stop_uoff = ~0ull; r.block_end = (uint32_t)f->blocks.size();
        size_t lo = 0; while (lo < f->blocks.size() && f->blocks[lo].uoff + f->blocks[lo].isize <= f->first_record_uoff) lo++;
        r.block_begin = (uint32_t)lo;
        if (r.block_begin < r.block_end) part.ranges.push_back(r);
        continue;
      }
      int ref = -1;
      for (size_t k = 0; k < f->ref_names.size(); k++) if (f->ref_names[k] == g.chrom) { ref = (int)k; break; }
      if (ref < 0) {
        // the reference surfaces this at execute time ("BAM region query failed" / "Reference not found in BAM header")
        ScanRange r = rule; r.region_mode = -1; part.ranges.push_back(r);
        continue;
      }
      const BaiRef* R = (size_t)ref < f->bai->refs.size() ? &f->bai->refs[(size_t)ref] : nullptr;
      if (g.unmapped_tail) {                                // physical_exec.rs:1131-1258
        rule.region_mode = 2; rule.region_ref = ref;
        uint64_t seek = 0;
        if (R) for (auto& kv : R->bins) for (auto& c : kv.second) seek = std::max(seek, c.end);
        if (seek == 0) {
          // last_first_record_start_position(): largest chunk start over the whole index; else the first record
          for (auto& RR : f->bai->refs) for (auto& kv : RR.bins) for (auto& c : kv.second) seek = std::max(seek, c.beg);
        }
        if (seek == 0) {
          ScanRange r = rule; r.exact_start = true; r.first_uoff = f->first_record_uoff; r.stop_uoff = ~0ull; r.block_end = (uint32_t)f->blocks.size();
          size_t lo = 0; while (lo < f->blocks.size() && f->blocks[lo].uoff + f->blocks[lo].isize <= f->first_record_uoff) lo++;
          r.block_begin = (uint32_t)lo;
          if (r.block_begin < r.block_end) part.ranges.push_back(r);
        } else { int rc = add_range(*f, seek, 0, rule, &part); if (rc) return rc; }
        continue;
      }
      rule.region_mode = 1; rule.region_ref = ref;
      rule.region_start = g.has_start ? g.start : 0; rule.region_end = g.has_end ? g.end : 0;
      if (!R) continue;
      if (g.has_start && g.has_end && g.start > g.end) continue;
      const bool q_has_end = g.has_end || has_start_sup;
      const uint64_t q_end = g.has_end ? (has_start_sup ? std::min(g.end, start_sup) : g.end) : start_sup;
      if (g.has_start && q_has_end && g.start > q_end) continue;
      std::vector<BaiChunk> chunks = bai_query(*f->bai, *R, g.has_start ? std::max<uint64_t>(g.start, 1) : 1, q_has_end, q_end);
      for (auto& c : chunks) { int rc = add_range(*f, c.beg, c.end, rule, &part); if (rc) return rc; }
    }
    plan->partitions.push_back(part);
  }
  *handled = true;
  return BAMSCAN_OK;
}

}  // namespace bamscan

// ---------------------------------------------------------------------------------------------
// ABI mirrors of the two pure planning functions, so their reference unit tests can be replayed against this library
using namespace bamscan;

extern "C" {

int bamscan_extract_regions(const BamScanFilter* filters, int32_t n_filters, int32_t zero_based, BamScanRegionAnalysis* out) {
  if (!out || (n_filters && !filters)) { set_error("null argument"); return BAMSCAN_ERR_INVALID; }
  RegionAnalysis A = analyze_regions(filters, n_filters, zero_based != 0);
  memset(out, 0, sizeof *out);
  out->unsatisfiable = A.unsatisfiable; out->has_start = A.has_lower; out->start = A.lower; out->has_end = A.has_upper; out->end = A.upper;
  size_t p = 0;
  for (auto& c : A.chroms) {
    if (p + c.size() + 1 > sizeof out->chroms) { set_error("too many chromosomes for BamScanRegionAnalysis"); return BAMSCAN_ERR_UNSUPPORTED; }
    memcpy(out->chroms + p, c.c_str(), c.size() + 1); p += c.size() + 1; out->n_chroms++;
  }
  for (int i = 0; i < n_filters && i < 64; i++) if (A.residual[(size_t)i]) out->residual_mask |= 1ull << i;
  return BAMSCAN_OK;
}

int bamscan_balance_partitions(const BamScanRegionEstimate* est, int32_t n, int32_t target_partitions, BamScanAssignedRegion* out,
                               int32_t cap, int32_t* n_out, int32_t* n_partitions) {
  if ((n && !est) || !n_out || !n_partitions) { set_error("null argument"); return BAMSCAN_ERR_INVALID; }
  std::vector<SizeEstimate> v;
  for (int i = 0; i < n; i++) {
    SizeEstimate e;
    e.region.chrom = est[i].chrom ? est[i].chrom : "";
    e.region.has_start = est[i].has_start; e.region.start = est[i].start; e.region.has_end = est[i].has_end; e.region.end = est[i].end;
    e.estimated_bytes = est[i].estimated_bytes; e.contig_length = est[i].contig_length; e.unmapped_count = est[i].unmapped_count;
    for (int k = 0; k < est[i].n_bin_positions; k++) e.nonempty_bin_positions.push_back(est[i].nonempty_bin_positions[k]);
    e.leaf_bin_span = est[i].leaf_bin_span; e.source_index = i;
    v.push_back(e);
  }
  std::vector<Assignment> a = balance_partitions(v, (size_t)std::max(0, target_partitions));
  int k = 0;
  for (size_t p = 0; p < a.size(); p++)
    for (size_t r = 0; r < a[p].regions.size(); r++, k++) {
      if (k >= cap) continue;
      const GenomicRegion& g = a[p].regions[r];
      out[k].partition = (int32_t)p; out[k].estimate_index = a[p].source[r];
      out[k].has_start = g.has_start; out[k].start = g.start; out[k].has_end = g.has_end; out[k].end = g.end;
      out[k].unmapped_tail = g.unmapped_tail; out[k].partition_total_estimated_bytes = a[p].total_estimated_bytes;
    }
  *n_out = k; *n_partitions = (int32_t)a.size();
  return BAMSCAN_OK;
}

}  // extern "C"
