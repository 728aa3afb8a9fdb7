// inflate_cta_core.h -- per-lane building blocks of the CTA-per-member inflate kernel (kernels_inflate_cta.cuh).
//
// Replaces (together with the kernel): noodles-bgzf 0.49.0 io::Reader block inflate + libdeflate `deflate_decompress`
// (reference call sites: datafusion/bio-format-bam/src/storage.rs:161-169, physical_exec.rs:409).
//
// Everything here is plain integer code over arrays (`pay` = the member's deflate payload as little-endian 32-bit words,
// `win` = the member's output window, `hb` = one "a parked match starts here" bit per window byte), written so that the SAME source is
// compiled into the kernel (arrays in shared memory, one call per thread) and into tools/inflate_sim.cpp, a host program
// that runs the lanes of a CTA in a loop -- the algorithm can be checked against zlib without a GPU.
//
// The algorithm (one CTA of NT lanes per BGZF member, all state in shared memory):
//   * a deflate block's symbol stream is cut into NT sub-streams at fixed bit positions; every lane decodes its own
//     sub-stream speculatively from the cut.  Huffman streams self-synchronise (measured on BAM data: half of the false
//     starts are on the true symbol chain after 6 symbols, 99 % after 46), so a lane's END position -- the first symbol
//     boundary at or after the next cut -- is almost always the true one even when its start was not;
//   * lanes then restart from their predecessor's end until nothing changes (lane 0 is exact, so lane i is exact after
//     at most i rounds; in practice two rounds).  This yields every lane's true start, its output byte count and the
//     lane holding the end-of-block symbol;
//   * an exclusive scan of the byte counts gives every lane its output offset; the lanes decode once more and EMIT:
//     literals go straight into the window, a match (len, dist) is parked in the first three bytes of its own
//     destination and its first byte is flagged in the head bitmap `hb`;
//   * RESOLVE (one warp, whole member, in output order): the head bitmap is read 32 words at a time, the matches of
//     the stripe are dealt to the lanes 32 at a time in position order, and a lane copies a piece (<= 16 bytes) of
//     its match as soon as the piece's source lies below the FRONTIER (the position of the first unfinished byte of
//     the batch: everything below it is final).  Measured on sorted BAM data the LZ77 dependency chains are ~120-190
//     matches deep per member (every record copies from the one before), i.e. ~33 independent matches per level:
//     one warp is the parallelism this stage has, and the frontier rule needs no flags, fences or polling.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define ICTA_HD __host__ __device__ __forceinline__
#else
#define ICTA_HD inline
#endif

namespace bamscan {
namespace icta {

constexpr int R_LL = 10, R_D = 8;                       // root LUT index bits (litlen / distance)
constexpr uint32_t ROOT_LL = 1u << R_LL, ROOT_D = 1u << R_D;
constexpr uint32_t SUB_LL = 512, SUB_D = 128;           // second-level entries available (a table that needs more is refused)

// 32-bit LUT entry (laid out for the fewest instructions in the decode step):
//   [4:0]   bits this half-token consumes: code length + extra bits (0 = unused code); E_SUB: index bits of the second level
//   [8:5]   code length (whole code, also in second-level entries)
//   [9] E_LIT literal   [10] E_SYM length / distance symbol   [11] E_EOB end of block   [12] E_SUB pointer to a second-level table
//   [31:16] literal byte | base length | base distance | E_SUB: index of the second-level table inside the LUT array
//   value = [31:16] + ((bits & mask([4:0])) >> [8:5])
constexpr uint32_t E_LIT = 1u << 9, E_SYM = 1u << 10, E_EOB = 1u << 11, E_SUB = 1u << 12, E_BAD = 0u;

enum Term : uint32_t { T_CROSS = 0, T_EOB = 1, T_BAD = 2 };    // how a lane's sub-stream decode ended
enum CoreErr : uint32_t { CE_OK = 0, CE_DIST = 5, CE_OVERRUN = 6 };   // values of InflateStatus (kernels_inflate.cuh)

ICTA_HD uint32_t entry_litlen(uint32_t sym, uint32_t len) {     // RFC 1951 3.2.5, computed arithmetically
  if (sym < 256u) return E_LIT | (len << 5) | len | (sym << 16);
  if (sym == 256u) return E_EOB | (len << 5) | len;
  const uint32_t s = sym - 257u;
  if (s > 28u) return E_BAD;
  const uint32_t extra = (s < 8u || s == 28u) ? 0u : (s - 4u) >> 2;
  const uint32_t base = s == 28u ? 258u : (s < 8u ? 3u + s : 3u + ((4u + (s & 3u)) << extra));
  return E_SYM | (len << 5) | (len + extra) | (base << 16);
}
ICTA_HD uint32_t entry_dist(uint32_t sym, uint32_t len) {
  if (sym > 29u) return E_BAD;
  const uint32_t extra = sym < 4u ? 0u : (sym - 2u) >> 1;
  const uint32_t base = sym < 4u ? 1u + sym : 1u + ((2u + (sym & 1u)) << extra);
  return E_SYM | (len << 5) | (len + extra) | (base << 16);
}
ICTA_HD uint32_t entry_sub(uint32_t root_bits, uint32_t sub_bits, uint32_t index) { return E_SUB | (root_bits << 5) | sub_bits | (index << 16); }

#if defined(__CUDA_ARCH__)
ICTA_HD void bm_or(uint32_t* p, uint32_t m) { atomicOr(p, m); }
ICTA_HD void bm_and(uint32_t* p, uint32_t m) { atomicAnd(p, m); }
ICTA_HD void fence_cta() { __threadfence_block(); }
ICTA_HD uint32_t ld_volatile(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
ICTA_HD uint32_t ffs32(uint32_t v) { return (uint32_t)__ffs((int)v); }
#else
ICTA_HD void bm_or(uint32_t* p, uint32_t m) { *p |= m; }
ICTA_HD void bm_and(uint32_t* p, uint32_t m) { *p &= m; }
ICTA_HD void fence_cta() {}
ICTA_HD uint32_t ld_volatile(const uint32_t* p) { return *p; }
ICTA_HD uint32_t ffs32(uint32_t v) { return (uint32_t)__builtin_ffs((int)v); }
#endif

struct SubResult { uint32_t end_bit, term, n_out, first_bit; };   // first_bit: first token boundary at or after count_from (0xffffffff: the chain ended before it)

// Decodes tokens from bit `start_bit` of `pay` until the first token boundary at or after `stop_bit`, an end-of-block
// symbol or an unused code.  EMIT = false: count only.  EMIT = true: literals and parked matches go into win[opos...]
// (obase = window offset of the member's first byte, olimit = window offset one past the member's last byte).
// A token is at most 48 bits; the caller guarantees that `pay` is readable 18 bytes past stop_bit.
//
// On the device ALL 32 lanes of a warp call this together (`enabled` = false for a lane with nothing to do): the loop
// condition is a warp vote, so the lanes re-converge after every token.  (A per-lane `while` with `continue` / `break`
// compiles to code whose lanes never re-converge inside the loop: measured, one active thread per issued instruction.)
#if !defined(__CUDACC__)
template <bool EMIT>
inline SubResult decode_sub(bool enabled, const uint32_t* pay, const uint32_t* lut_ll, const uint32_t* lut_d,
                            uint32_t start_bit, uint32_t count_from, uint32_t stop_bit,
                            uint8_t* win, uint32_t* hb, uint32_t opos, uint32_t obase, uint32_t olimit, uint32_t* err) {
  uint32_t wi = enabled ? (start_bit >> 5) : 0u;
  const uint32_t sh = start_bit & 31u;
  uint64_t buf = (((uint64_t)pay[wi + 1] << 32) | pay[wi]) >> sh;
  uint32_t cnt = 64u - sh;                                   // valid bits in buf
  wi += 2;
  uint32_t pos = start_bit, n_out = 0, term = T_CROSS, first_bit = 0xffffffffu;
  bool live = enabled && pos < stop_bit;
  while (live) {
    const bool counting = pos >= count_from;                 // tokens in front of count_from only warm the chain up
    if (counting && first_bit == 0xffffffffu) first_bit = pos;
    if (cnt <= 32u) { buf |= (uint64_t)pay[wi] << cnt; cnt += 32u; wi++; }
    uint32_t bits = (uint32_t)buf;
    uint32_t e = lut_ll[bits & (ROOT_LL - 1u)];
    if (e & E_SUB) e = lut_ll[(e >> 16) + ((bits >> R_LL) & ~(~0u << (e & 31u)))];
    uint32_t tot = e & 31u;
    const uint32_t len = (e >> 16) + ((bits & ~(~0u << tot)) >> ((e >> 5) & 15u));   // literal: its byte
    buf >>= tot; cnt -= tot; pos += tot;
    if (e & E_LIT) {
      if (EMIT) {
        if (opos >= olimit) { *err = CE_OVERRUN; term = T_BAD; }
        else { win[opos] = (uint8_t)len; opos++; }
      }
      if (counting) n_out++;
    } else if (e & E_SYM) {
      if (cnt <= 32u) { buf |= (uint64_t)pay[wi] << cnt; cnt += 32u; wi++; }
      bits = (uint32_t)buf;
      uint32_t de = lut_d[bits & (ROOT_D - 1u)];
      if (de & E_SUB) de = lut_d[(de >> 16) + ((bits >> R_D) & ~(~0u << (de & 31u)))];
      if (!(de & E_SYM)) term = T_BAD;
      else {
        tot = de & 31u;
        const uint32_t dist = (de >> 16) + ((bits & ~(~0u << tot)) >> ((de >> 5) & 15u));
        buf >>= tot; cnt -= tot; pos += tot;
        if (EMIT) {
          if (dist > opos - obase) { *err = CE_DIST; term = T_BAD; }
          else if (opos + len > olimit) { *err = CE_OVERRUN; term = T_BAD; }
          else {
            const uint32_t v = (dist - 1u) | ((len - 3u) << 15);   // parked in the match's own first three bytes
            win[opos] = (uint8_t)v; win[opos + 1] = (uint8_t)(v >> 8); win[opos + 2] = (uint8_t)(v >> 16);
            hb[opos >> 5] |= 1u << (opos & 31u);                    // head bit
            opos += len;
          }
        }
        if (counting) n_out += len;
      }
    } else {
      term = (e & E_EOB) ? T_EOB : T_BAD;
    }
    live = term == T_CROSS && pos < stop_bit;
  }
  if (first_bit == 0xffffffffu && term == T_CROSS && pos >= count_from) first_bit = pos;
  SubResult r; r.end_bit = pos; r.term = term; r.n_out = n_out; r.first_bit = first_bit;
  return r;
}
#else
// Device version: same contract, shared-memory state addresses (32-bit shared-window addresses), and ALL 32 lanes of a warp
// call it together (`enabled` = false for a lane with nothing to do).  The loop condition is a warp vote, so the lanes
// re-converge after every token; the body is straight-line predicated code: literal/length half, then the distance half,
// which every lane executes (a literal lane discards it).  A per-lane `while` with `continue` / `break` compiles to code
// whose lanes never re-converge inside the loop (measured: one active thread per issued instruction).
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
template <bool EMIT>
__device__ __forceinline__ SubResult decode_sub(bool enabled, uint32_t pay_s, uint32_t ll_s, uint32_t d_s,
                                                uint32_t start_bit, uint32_t count_from, uint32_t stop_bit,
                                                uint32_t win_s, uint32_t hb_s, uint32_t opos, uint32_t obase, uint32_t olimit, uint32_t* err) {
  uint32_t wa = pay_s + ((enabled ? start_bit : 0u) >> 5) * 4u;
  const uint32_t sh = start_bit & 31u;
  uint64_t buf = (((uint64_t)lds_u32(wa + 4u) << 32) | lds_u32(wa)) >> sh;
  uint32_t cnt = 64u - sh;
  wa += 8u;
  uint32_t pos = start_bit, n_out = 0, term = T_CROSS, first_bit = 0xffffffffu;
  bool live = enabled && pos < stop_bit;
  while (__any_sync(0xffffffffu, live)) {
    const bool counting = live && pos >= count_from;           // tokens in front of count_from only warm the chain up
    first_bit = min(first_bit, counting ? pos : 0xffffffffu);
    if (cnt <= 32u) { buf |= (uint64_t)lds_u32(wa) << cnt; cnt += 32u; wa += 4u; }     // (a finished lane refills at most once more)
    uint32_t bits = (uint32_t)buf;
    uint32_t e = lds_u32(ll_s + ((bits & (ROOT_LL - 1u)) << 2));
    if (e & E_SUB) e = lds_u32(ll_s + (((e >> 16) + ((bits >> R_LL) & ~(~0u << (e & 31u)))) << 2));
    uint32_t tot = live ? (e & 31u) : 0u;
    const uint32_t len = (e >> 16) + ((bits & ~(~0u << tot)) >> ((e >> 5) & 15u));
    const bool lit = live && (e & E_LIT) != 0, mat = live && (e & E_SYM) != 0;
    buf >>= tot; cnt -= tot; pos += tot;
    if (live && !(e & (E_LIT | E_SYM))) term = (e & E_EOB) ? T_EOB : T_BAD;
    if (lit) {
      if (EMIT) {
        if (opos >= olimit) { *err = CE_OVERRUN; term = T_BAD; }
        else { sts_u8(win_s + opos, len); opos++; }
      }
      n_out += counting ? 1u : 0u;
    }
    if (cnt <= 32u) { buf |= (uint64_t)lds_u32(wa) << cnt; cnt += 32u; wa += 4u; }
    bits = (uint32_t)buf;
    uint32_t de = lds_u32(d_s + ((bits & (ROOT_D - 1u)) << 2));
    if (de & E_SUB) de = lds_u32(d_s + (((de >> 16) + ((bits >> R_D) & ~(~0u << (de & 31u)))) << 2));
    const bool dok = mat && (de & E_SYM) != 0;
    if (mat && !dok) term = T_BAD;
    tot = dok ? (de & 31u) : 0u;
    if (EMIT) {
      if (dok) {
        const uint32_t dist = (de >> 16) + ((bits & ~(~0u << tot)) >> ((de >> 5) & 15u));
        if (dist > opos - obase) { *err = CE_DIST; term = T_BAD; }
        else if (opos + len > olimit) { *err = CE_OVERRUN; term = T_BAD; }
        else {
          const uint32_t v = (dist - 1u) | ((len - 3u) << 15);     // parked in the match's own first three bytes
          sts_u8(win_s + opos, v); sts_u8(win_s + opos + 1u, v >> 8); sts_u8(win_s + opos + 2u, v >> 16);
          asm volatile("atom.shared.or.b32 _, [%0], %1;" ::"r"(hb_s + ((opos >> 5) << 2)), "r"(1u << (opos & 31u)) : "memory");   // head bit
          opos += len;
        }
      }
    }
    buf >>= tot; cnt -= tot; pos += tot;
    if (dok && counting) n_out += len;
    live = live && term == T_CROSS && pos < stop_bit;
  }
  if (first_bit == 0xffffffffu && term == T_CROSS && pos >= count_from) first_bit = pos;
  SubResult r; r.end_bit = pos; r.term = term; r.n_out = n_out; r.first_bit = first_bit;
  return r;
}
#endif

constexpr uint32_t RESOLVE_PIECE = 16;   // bytes a lane copies per round of the resolver (longer matches continue in the next round)

}  // namespace icta
}  // namespace bamscan
