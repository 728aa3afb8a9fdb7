// kernels_inflate.cuh -- block-parallel raw-DEFLATE (RFC 1951) inflate + CRC-32 for BGZF members.
//
// Replaces: noodles-bgzf 0.49.0 io::Reader block inflate + libdeflate `deflate_decompress`/`crc32`
// (reference call sites: datafusion/bio-format-bam/src/storage.rs:161-169, physical_exec.rs:409).
//
// Mapping to the machine (sm_100a):
//   * one WARP decodes one BGZF member (<= 64 KiB out); warps pull members from a global atomic
//     ticket so a persistent grid of 148 x CTAS_PER_SM CTAs stays balanced;
//   * Huffman state lives in shared memory, ~3.6 KB per warp: 10-bit litlen LUT + 8-bit distance
//     LUT (16-bit entries) + canonical (sorted-symbol) arrays for the rare codes longer than the
//     LUT index -> 48 resident warps / SM;
//   * tables are built warp-cooperatively (match_any ranks, brev canonical codes, strided fill);
//   * the symbol loop is warp-UNIFORM (no divergence): the bit reader is two 32-bit registers plus
//     a 128-byte register window (one word per lane, double buffered) refilled by shuffles, so the
//     compressed stream is read from HBM exactly once, in full 128-byte lines;
//   * literals are parked one per lane and flushed as 32-byte coalesced stores; LZ77 copies are
//     done by all 32 lanes (dist==1 broadcast, dist<32 modular, else strided with warp fences);
//   * CRC-32 of the member is computed by the same warp while the bytes are still in L1/L2:
//     32 lane-chunks by slicing-by-4 on shared tables, then combined with x^(8k) mod P multiplies.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bamscan {

struct BlockDesc {        // one BGZF member of the chunk (host-built from the BSIZE chain)
  uint32_t cdata_off;     // byte offset of the raw-deflate payload inside the chunk's compressed buffer
  uint32_t cdata_len;     // payload bytes (BSIZE+1 - header - 8)
  uint32_t isize;         // expected inflated size (trailer ISIZE)
  uint32_t crc;           // expected CRC-32 (trailer)
  uint32_t uoff;          // offset of this member's output inside the chunk's inflated buffer
};

enum InflateStatus : uint32_t {
  INF_OK = 0, INF_ERR_BTYPE = 1, INF_ERR_STORED = 2, INF_ERR_TABLE = 3, INF_ERR_SYMBOL = 4,
  INF_ERR_DIST = 5, INF_ERR_OVERRUN = 6, INF_ERR_ISIZE = 7, INF_ERR_CRC = 8, INF_ERR_INPUT = 9
};

constexpr int INF_WARPS = 8;                 // warps per CTA
constexpr int INF_CTAS_PER_SM = 6;
constexpr int INF_LL_BITS = 9, INF_D_BITS = 7;
constexpr uint32_t FULL = 0xffffffffu;

// 32-bit LUT entry, decoded without table look-ups or branches:
//   [3:0]  code length (0 = invalid code)      [7:4] number of extra bits
//   [10:8] kind bits: 0x100 length / distance, 0x200 end of block, 0x400 code longer than the LUT index (canonical walk) or invalid
//   [30:16] literal byte | base length | base distance        [31] literal flag (sign test on the hot path)
// value = base + ((bits >> code_len) & mask(extra)) ; bits consumed = code_len + extra
constexpr uint32_t E_LIT = 0x80000000u, E_SYM = 1u << 8, E_EOB = 2u << 8, E_LONG = 4u << 8;   // one bit each: single-instruction tests

struct WarpTables {
  uint32_t lut_ll[1 << INF_LL_BITS];
  uint32_t lut_d[1 << INF_D_BITS];
  uint16_t sorted_ll[288];
  uint16_t sorted_d[32];
  uint16_t first_ll[16], offs_ll[16], cnt_ll[16];
  uint16_t first_d[16], offs_d[16], cnt_d[16];
  uint16_t nxt[16];
  uint8_t cl[320];          // code lengths being assembled (litlen then dist)
};

struct InflateShared {
  uint32_t crc_tab[4][256];
  WarpTables wt[INF_WARPS];
};

__constant__ uint8_t c_clc_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
// x^(8 * 2^j) mod P (reflected CRC-32 domain), j = 0..17 ; filled by the host at init
__constant__ uint32_t c_crc_xpow8[18];

enum TableKind { TK_LITLEN = 0, TK_DIST = 1, TK_PRECODE = 2 };

// LUT entry of symbol `sym` with code length `len` (RFC 1951 3.2.5 base/extra tables, computed arithmetically)
template <int TK>
__device__ __forceinline__ uint32_t make_entry(uint32_t sym, uint32_t len) {
  if (TK == TK_PRECODE) return E_SYM | len | (sym << 16);
  if (TK == TK_LITLEN) {
    if (sym < 256u) return E_LIT | len | (sym << 16);
    if (sym == 256u) return E_EOB | len;
    uint32_t s = sym - 257u;
    if (s > 28u) return E_LONG;                                   // 286, 287: invalid
    uint32_t extra = (s < 8u || s == 28u) ? 0u : (s - 4u) >> 2;
    uint32_t base = s == 28u ? 258u : (s < 8u ? 3u + s : 3u + ((4u + (s & 3u)) << extra));
    return E_SYM | len | (extra << 4) | (base << 16);
  }
  if (sym > 29u) return E_LONG;
  uint32_t extra = sym < 4u ? 0u : (sym - 2u) >> 1;
  uint32_t base = sym < 4u ? 1u + sym : 1u + ((2u + (sym & 1u)) << extra);
  return E_SYM | len | (extra << 4) | (base << 16);
}

// ---------------------------------------------------------------------------------------------
// warp-uniform bit reader: every lane holds the same 64-bit window (lo:hi) plus one prefetched word.  Refills are
// plain warp-uniform loads (one 32-byte sector, L1-resident after the first touch of a line), issued one word ahead
// so their latency is off the critical path.  Invariant after init()/consume(): bp < 32 => peek() returns 32 valid bits.
struct BitReader {
  const uint32_t* base;   // 4-byte aligned start
  uint32_t lo, hi, nxt;   // words wi-3, wi-2, wi-1
  uint32_t bp;            // 0..31: position of the next unread bit inside lo
  uint32_t wi;            // index of the next word to load
  uint32_t limit_words;   // words that may legitimately be touched (+ slack)

  __device__ __forceinline__ void init(const uint8_t* p, uint32_t nbytes) {
    uintptr_t a = reinterpret_cast<uintptr_t>(p);
    base = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
    bp = uint32_t(a & 3) * 8;
    limit_words = (uint32_t(a & 3) + nbytes + 3) / 4 + 6;
    lo = __ldg(base); hi = __ldg(base + 1); nxt = __ldg(base + 2);
    wi = 3;
  }
  __device__ __forceinline__ uint32_t peek() const { return __funnelshift_r(lo, hi, bp); }
  __device__ __forceinline__ void consume(uint32_t n) {   // n <= 32
    bp += n;
    if (bp >= 32u) { lo = hi; hi = nxt; nxt = __ldg(base + wi); wi++; bp -= 32u; }
  }
  __device__ __forceinline__ uint32_t take(uint32_t n) {   // n < 32
    uint32_t v = peek() & ((1u << n) - 1u);
    consume(n);
    return v;
  }
  __device__ __forceinline__ bool exhausted() const { return wi > limit_words; }
  // byte address of the next unread bit (must be byte aligned)
  __device__ __forceinline__ const uint8_t* byte_ptr() const {
    return reinterpret_cast<const uint8_t*>(base + (wi - 3)) + (bp >> 3);
  }
};

// ---------------------------------------------------------------------------------------------
// Canonical Huffman table build, warp-cooperative.  cl[0..n) code lengths (0 = unused).
// Returns false when the code is over-subscribed.
template <int PBITS, int TK>
__device__ __forceinline__ bool build_table(const uint8_t* cl, int n, uint32_t* lut, uint16_t* sorted,
                                            uint16_t* first, uint16_t* offs, uint16_t* cnt, uint16_t* nxt, int lane) {
  if (lane < 16) { cnt[lane] = 0; nxt[lane] = 0; }
  for (int i = lane; i < (1 << PBITS); i += 32) lut[i] = E_LONG;          // kind LONG with length 0 = invalid code
  __syncwarp();
  for (int s = lane; s < n; s += 32) {
    uint32_t L = cl[s];
    if (L) atomicAdd(reinterpret_cast<unsigned int*>(cnt) + (L >> 1), (L & 1) ? 0x10000u : 1u);   // 16-bit counters packed in pairs
  }
  __syncwarp();
  // every lane derives first codes / offsets redundantly; lane == len stores its row
  uint32_t code = 0, off = 0, left = 1;
  bool over = false;
  for (int len = 1; len <= 15; len++) {
    uint32_t c = cnt[len];
    code = (code + cnt[len - 1]) << 1;
    left <<= 1;
    if (c > left) over = true;
    left -= c;
    if (lane == len) { first[len] = (uint16_t)code; offs[len] = (uint16_t)off; }
    off += c;
  }
  if (over) return false;
  __syncwarp();
  const uint32_t lt = (1u << lane) - 1u;
  for (int b = 0; b < n; b += 32) {
    int s = b + lane;
    uint32_t L = (s < n) ? cl[s] : 0u;
    uint32_t m = __match_any_sync(FULL, L);
    uint32_t r = nxt[L] + __popc(m & lt);
    __syncwarp();
    if (L && (m & lt) == 0) nxt[L] = (uint16_t)(nxt[L] + __popc(m));
    __syncwarp();
    if (L) {
      sorted[offs[L] + r] = (uint16_t)s;
      uint32_t cd = first[L] + r;
      uint32_t rev = __brev(cd) >> (32 - L);
      if (L <= (uint32_t)PBITS) {
        uint32_t e = make_entry<TK>((uint32_t)s, L);
        for (uint32_t idx = rev; idx < (1u << PBITS); idx += (1u << L)) lut[idx] = e;
      } else {
        lut[rev & ((1u << PBITS) - 1u)] = E_LONG | 1u;   // long-code marker (length field != 0)
      }
    }
  }
  __syncwarp();
  return true;
}

// Canonical walk for codes longer than the LUT index.  Returns E_LONG (length 0) for an invalid code.
template <int PBITS, int TK>
__device__ __noinline__ uint32_t decode_long(uint32_t bits, const uint16_t* sorted, const uint16_t* first, const uint16_t* offs, const uint16_t* cnt) {
  uint32_t rb = __brev(bits);
  for (uint32_t len = PBITS + 1; len <= 15; len++) {
    uint32_t d = (rb >> (32 - len)) - first[len];
    if (d < cnt[len]) return make_entry<TK>(sorted[offs[len] + d], len);
  }
  return E_LONG;
}

__device__ __forceinline__ uint32_t crc_mulmod(uint32_t a, uint32_t b) {   // a*b mod P, reflected domain
  uint32_t p = 0;
  #pragma unroll 4
  for (int i = 0; i < 32; i++) {
    p ^= (b & 0x80000000u) ? a : 0u;
    a = (a >> 1) ^ ((a & 1u) ? 0xEDB88320u : 0u);
    b <<= 1;
  }
  return p;
}
__device__ __forceinline__ uint32_t crc_shift(uint32_t crc, uint32_t nbytes) {   // crc * x^(8*nbytes)
  for (int j = 0; nbytes; j++, nbytes >>= 1)
    if (nbytes & 1u) crc = crc_mulmod(crc, c_crc_xpow8[j]);
  return crc;
}

// CRC-32 (IEEE, reflected, init/final 0xffffffff) of out[0..n) by one warp.
__device__ __noinline__ uint32_t warp_crc32(const uint8_t* out, uint32_t n, const uint32_t (*tab)[256], int lane) {
  uint32_t chunk = ((n + 31) / 32 + 3) & ~3u;
  uint32_t b = min(n, chunk * lane), e = min(n, b + chunk);
  uint32_t st = (lane == 0) ? 0xffffffffu : 0u;
  const uint8_t* p = out + b;
  uint32_t len = e - b;
  while (len && (reinterpret_cast<uintptr_t>(p) & 3)) { st = tab[0][(st ^ *p) & 0xff] ^ (st >> 8); p++; len--; }
  const uint32_t* pw = reinterpret_cast<const uint32_t*>(p);
  for (; len >= 4; len -= 4) {
    st ^= *pw++;
    st = tab[3][st & 0xff] ^ tab[2][(st >> 8) & 0xff] ^ tab[1][(st >> 16) & 0xff] ^ tab[0][st >> 24];
  }
  p = reinterpret_cast<const uint8_t*>(pw);
  while (len) { st = tab[0][(st ^ *p) & 0xff] ^ (st >> 8); p++; len--; }
  st = crc_shift(st, n - e);
  #pragma unroll
  for (int o = 16; o; o >>= 1) st ^= __shfl_xor_sync(FULL, st, o);
  return ~st;
}

// Reads the dynamic-block header (HLIT/HDIST/HCLEN + code lengths) into T.cl.  Returns an InflateStatus.
__device__ __forceinline__ uint32_t read_dynamic_header(BitReader& br, WarpTables& T, int lane, int* n_ll_out, int* n_d_out) {
  uint32_t h = br.take(14);
  int n_ll = (int)(h & 31u) + 257, n_d = (int)((h >> 5) & 31u) + 1;
  int n_clc = (int)(h >> 10) + 4;
  if (n_ll > 286 || n_d > 30) return INF_ERR_TABLE;
  if (lane < 19) T.cl[lane] = 0;
  __syncwarp();
  for (int i = 0; i < n_clc; i++) {
    uint32_t v = br.take(3);
    if (lane == 0) T.cl[c_clc_order[i]] = (uint8_t)v;
  }
  __syncwarp();
  // pre-code (max length 7) goes into lut_d; the decoded lengths then overwrite cl[]
  if (!build_table<7, TK_PRECODE>(T.cl, 19, T.lut_d, T.sorted_d, T.first_d, T.offs_d, T.cnt_d, T.nxt, lane)) return INF_ERR_TABLE;
  int total = n_ll + n_d, i = 0;
  uint32_t prev = 0;
  while (i < total) {
    uint32_t bits = br.peek();
    uint32_t e = T.lut_d[bits & 127u];
    uint32_t L = e & 15u, s = e >> 16;
    if (L == 0) return INF_ERR_TABLE;
    uint32_t rep, val, xb;
    if (s < 16) { rep = 1; val = s; prev = s; xb = 0; }
    else if (s == 16) { if (i == 0) return INF_ERR_TABLE; xb = 2; rep = 3 + ((bits >> L) & 3u); val = prev; }
    else if (s == 17) { xb = 3; rep = 3 + ((bits >> L) & 7u); val = 0; prev = 0; }
    else { xb = 7; rep = 11 + ((bits >> L) & 127u); val = 0; prev = 0; }
    br.consume(L + xb);
    if (i + (int)rep > total) return INF_ERR_TABLE;
    for (uint32_t k = lane; k < rep; k += 32) T.cl[i + k] = (uint8_t)val;
    i += (int)rep;
  }
  __syncwarp();
  if (T.cl[256] == 0) return INF_ERR_TABLE;
  *n_ll_out = n_ll; *n_d_out = n_d;
  return INF_OK;
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(INF_WARPS * 32, INF_CTAS_PER_SM)
inflate_kernel(const uint8_t* __restrict__ comp, const BlockDesc* __restrict__ blocks, uint32_t n_blocks,
               uint8_t* __restrict__ infl, uint32_t* __restrict__ status, uint32_t* __restrict__ ticket,
               uint32_t* __restrict__ err_flag, int check_crc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  InflateShared& sh = *reinterpret_cast<InflateShared*>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // CRC tables (slicing-by-4)
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    uint32_t c = (uint32_t)i;
    for (int k = 0; k < 8; k++) c = (c >> 1) ^ ((c & 1u) ? 0xEDB88320u : 0u);
    sh.crc_tab[0][i] = c;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    uint32_t c = sh.crc_tab[0][i];
    for (int t = 1; t < 4; t++) { c = sh.crc_tab[0][c & 0xff] ^ (c >> 8); sh.crc_tab[t][i] = c; }
  }
  __syncthreads();
  WarpTables& T = sh.wt[warp];
  const bool lane0 = lane == 0;

  for (;;) {
    uint32_t bi = 0;
    if (lane0) bi = atomicAdd(ticket, 1u);
    bi = __shfl_sync(FULL, bi, 0);
    if (bi >= n_blocks) break;
    const BlockDesc bd = blocks[bi];
    uint8_t* const out = infl + bd.uoff;
    const uint32_t isize = bd.isize;
    uint32_t outpos = 0, err = INF_OK;
    BitReader br;
    br.init(comp + bd.cdata_off, bd.cdata_len);
    const uint32_t obase = bd.uoff;                 // infl[obase + k]: uniform base + 32-bit offset addressing
    bool final_block = false;

    while (!final_block && err == INF_OK) {
      if (br.exhausted()) { err = INF_ERR_INPUT; break; }   // every path below consumes input, so this bounds the loop
      uint32_t hdr = br.take(3);
      final_block = hdr & 1u;
      uint32_t btype = hdr >> 1;
      if (btype == 0) {
        // stored: skip to byte boundary, LEN / NLEN, raw copy
        br.consume((8 - (br.bp & 7)) & 7);
        uint32_t len = br.take(16), nlen = br.take(16);
        if ((len ^ nlen) != 0xffffu || outpos + len > isize) { err = INF_ERR_STORED; break; }
        const uint8_t* src = br.byte_ptr();
        uint32_t used = (uint32_t)(src + len - (comp + bd.cdata_off));
        if (used > bd.cdata_len) { err = INF_ERR_INPUT; break; }
        for (uint32_t i = lane; i < len; i += 32) out[outpos + i] = src[i];
        outpos += len;
        br.init(src + len, bd.cdata_len - used);
        continue;
      }
      if (btype == 3) { err = INF_ERR_BTYPE; break; }
      int n_ll = 288, n_d = 30;
      if (btype == 1) {
        for (int i = lane; i < 288; i += 32) T.cl[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
        if (lane < 30) T.cl[288 + lane] = 5;
        __syncwarp();
      } else {
        err = read_dynamic_header(br, T, lane, &n_ll, &n_d);
        if (err) break;
      }
      if (!build_table<INF_LL_BITS, TK_LITLEN>(T.cl, n_ll, T.lut_ll, T.sorted_ll, T.first_ll, T.offs_ll, T.cnt_ll, T.nxt, lane)) { err = INF_ERR_TABLE; break; }
      if (!build_table<INF_D_BITS, TK_DIST>(T.cl + n_ll, n_d, T.lut_d, T.sorted_d, T.first_d, T.offs_d, T.cnt_d, T.nxt, lane)) { err = INF_ERR_TABLE; break; }

      // ---------------- symbol loop (warp-uniform; one look-up decodes code + extra bits) ----------------
      // The reader state lives in plain registers here; refills are real (rare) branches so the common path stays short.
      // Output overrun is checked at every match and at the end of the member (the inflated buffer carries slack for
      // literal runs); the loop is bounded by the input: the refill path stops at the member's last word (+ slack).
      {
        uint32_t lo = br.lo, hi = br.hi, nxt = br.nxt, bp = br.bp, wi = br.wi;
        const uint32_t* const wbase = br.base;
        const uint32_t wlimit = br.limit_words;
        const uint32_t* const lut_ll = T.lut_ll;
        const uint32_t* const lut_d = T.lut_d;
        uint8_t* outl = infl + obase + outpos + lane;        // this lane's byte of the next output stripe
#define INF_REFILL()                                                                         \
        if (bp >= 32u) {                                                                     \
          lo = hi; hi = nxt; nxt = __ldg(wbase + wi); wi++; bp -= 32u;                        \
          if (wi > wlimit) { err = INF_ERR_INPUT; break; }                                    \
        }
        for (;;) {
          uint32_t bits = __funnelshift_r(lo, hi, bp);
          uint32_t e = lut_ll[bits & ((1u << INF_LL_BITS) - 1u)];
          if ((int32_t)e < 0) {                                // literal
            if (lane0) *outl = (uint8_t)(e >> 16);
            outl++; outpos++;
            bp += e & 15u;
            INF_REFILL();
            continue;
          }
          if (!(e & E_SYM)) {                                  // rare: code longer than the LUT index, end of block, invalid code
            if (e & E_LONG) {
              e = decode_long<INF_LL_BITS, TK_LITLEN>(bits, T.sorted_ll, T.first_ll, T.offs_ll, T.cnt_ll);
              if ((int32_t)e < 0) {
                if (lane0) *outl = (uint8_t)(e >> 16);
                outl++; outpos++;
                bp += e & 15u;
                INF_REFILL();
                continue;
              }
            }
            if (!(e & E_SYM)) {
              bp += e & 15u;
              if ((e & 15u) == 0) err = INF_ERR_SYMBOL;
              break;
            }
          }
          {
            const uint32_t nb = e & 15u, xb = (e >> 4) & 15u;
            const uint32_t len = (e >> 16) + ((bits >> nb) & ((1u << xb) - 1u));
            bp += nb + xb;
            INF_REFILL();
            bits = __funnelshift_r(lo, hi, bp);
            uint32_t de = lut_d[bits & ((1u << INF_D_BITS) - 1u)];
            if (de & E_LONG) de = decode_long<INF_D_BITS, TK_DIST>(bits, T.sorted_d, T.first_d, T.offs_d, T.cnt_d);
            const uint32_t dnb = de & 15u, dxb = (de >> 4) & 15u;
            const uint32_t dist = (de >> 16) + ((bits >> dnb) & ((1u << dxb) - 1u));
            bp += dnb + dxb;
            if (dnb == 0 || dist > outpos || outpos + len > isize) { err = (dnb == 0 || dist > outpos) ? INF_ERR_DIST : INF_ERR_OVERRUN; break; }
            __syncwarp();                                   // earlier stores (other lanes) become visible to the loads below
            // byte i of the match is src[i mod dist]: only bytes that already exist are read, whatever dist/len are
            if (dist >= len && len <= 32u) {                // the common case: one stripe, no overlap
              const uint8_t v = *(outl - dist);
              if ((uint32_t)lane < len) *outl = v;
            } else {
              uint8_t* const d0 = outl - lane;
              const uint8_t* const s0 = d0 - dist;
              if (dist >= len) { for (uint32_t i = lane; i < len; i += 32) d0[i] = s0[i]; }
              else if (dist == 1) { const uint8_t v = s0[0]; for (uint32_t i = lane; i < len; i += 32) d0[i] = v; }
              else { for (uint32_t i = lane; i < len; i += 32) d0[i] = s0[i % dist]; }
            }
            outl += len; outpos += len;
            INF_REFILL();
          }
        }
#undef INF_REFILL
        br.lo = lo; br.hi = hi; br.nxt = nxt; br.bp = bp; br.wi = wi;
        // normalise (a block may end with bp in [32, 47))
        if (br.bp >= 32u) { br.lo = br.hi; br.hi = br.nxt; br.nxt = __ldg(br.base + br.wi); br.wi++; br.bp -= 32u; }
      }
      if (br.exhausted() && err == INF_OK) err = INF_ERR_INPUT;
    }
    __syncwarp();
    if (err == INF_OK && outpos != isize) err = outpos > isize ? INF_ERR_OVERRUN : INF_ERR_ISIZE;
    if (err == INF_OK && check_crc) {
      uint32_t crc = warp_crc32(out, isize, sh.crc_tab, lane);
      if (crc != bd.crc) err = INF_ERR_CRC;
    }
    if (lane0) {
      status[bi] = err;
      if (err) atomicCAS(err_flag, 0u, (bi << 4) | err | 0x80000000u);
    }
    __syncwarp();
  }
}

}  // namespace bamscan
