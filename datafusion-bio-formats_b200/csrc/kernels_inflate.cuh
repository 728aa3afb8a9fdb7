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
constexpr int INF_LL_BITS = 10, INF_D_BITS = 8;
constexpr uint32_t FULL = 0xffffffffu;

// 16-bit LUT entry: [15:14] kind, [13:10] code length, [9:0] value
//   kind 0 literal (value = byte) | 1 length symbol (value = sym-257) / distance symbol | 2 end of block | 3 long code / invalid (len 0 = invalid)
constexpr uint32_t E_LIT = 0u << 14, E_SYM = 1u << 14, E_EOB = 2u << 14, E_LONG = 3u << 14;

struct WarpTables {
  uint16_t lut_ll[1 << INF_LL_BITS];
  uint16_t lut_d[1 << INF_D_BITS];
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

__constant__ uint16_t c_len_base[32] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258, 0, 0, 0};
__constant__ uint8_t c_len_extra[32] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0, 0, 0, 0};
__constant__ uint16_t c_dist_base[32] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577, 0, 0};
__constant__ uint8_t c_dist_extra[32] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13, 0, 0};
__constant__ uint8_t c_clc_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
// x^(8 * 2^j) mod P (reflected CRC-32 domain), j = 0..17 ; filled by the host at init
__constant__ uint32_t c_crc_xpow8[18];

// ---------------------------------------------------------------------------------------------
// warp-uniform bit reader over a double-buffered 128-byte register window
struct BitReader {
  const uint32_t* base;   // 4-byte aligned start
  uint32_t win, winnext;  // lane's word of the current / next 32-word window
  uint32_t lo, hi;        // 64 valid bits starting at bit `bp` of lo
  uint32_t bp;            // 0..31
  uint32_t wi;            // index of the next word to fetch
  uint32_t limit_words;   // words that may legitimately be consumed (+ slack)
  bool overrun;

  __device__ __forceinline__ void init(const uint8_t* p, uint32_t nbytes, int lane) {
    uintptr_t a = reinterpret_cast<uintptr_t>(p);
    base = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
    bp = uint32_t(a & 3) * 8;
    limit_words = (uint32_t(a & 3) + nbytes + 3) / 4 + 2;
    win = __ldg(base + lane);
    winnext = __ldg(base + 32 + lane);
    lo = __shfl_sync(FULL, win, 0);
    hi = __shfl_sync(FULL, win, 1);
    wi = 2;
    overrun = false;
  }
  __device__ __forceinline__ uint32_t peek() const { return __funnelshift_r(lo, hi, bp); }
  __device__ __forceinline__ void consume(uint32_t n, int lane) {
    bp += n;
    if (bp >= 32) {
      bp -= 32;
      lo = hi;
      hi = __shfl_sync(FULL, win, wi & 31);
      wi++;
      if ((wi & 31) == 0) {
        win = winnext;
        if (wi > limit_words) overrun = true;
        winnext = __ldg(base + wi + 32 + lane);
      }
    }
  }
  __device__ __forceinline__ uint32_t take(uint32_t n, int lane) {   // n <= 24... (any n < 32)
    uint32_t v = peek() & ((1u << n) - 1u);
    consume(n, lane);
    return v;
  }
  // byte address of the next unread bit (must be byte aligned)
  __device__ __forceinline__ const uint8_t* byte_ptr() const {
    return reinterpret_cast<const uint8_t*>(base + (wi - 2)) + (bp >> 3);
  }
};

// ---------------------------------------------------------------------------------------------
// Canonical Huffman table build, warp-cooperative.  cl[0..n) code lengths (0 = unused).
// Returns false when the code is over-subscribed.
template <int PBITS, bool IS_LITLEN>
__device__ __forceinline__ bool build_table(const uint8_t* cl, int n, uint16_t* lut, uint16_t* sorted,
                                            uint16_t* first, uint16_t* offs, uint16_t* cnt, uint16_t* nxt, int lane) {
  if (lane < 16) { cnt[lane] = 0; nxt[lane] = 0; }
  // invalidate LUT (kind LONG with len 0 = invalid)
  for (int i = lane; i < (1 << PBITS) / 2; i += 32) reinterpret_cast<uint32_t*>(lut)[i] = (E_LONG << 16) | E_LONG;
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
    code = (code + (len > 1 ? cnt[len - 1] : 0u)) << 1;
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
        uint32_t e;
        if (IS_LITLEN) e = s < 256 ? (E_LIT | (uint32_t)s) : (s == 256 ? E_EOB : (E_SYM | (uint32_t)(s - 257)));
        else e = E_SYM | (uint32_t)s;
        e |= L << 10;
        for (uint32_t idx = rev; idx < (1u << PBITS); idx += (1u << L)) lut[idx] = (uint16_t)e;
      } else {
        lut[rev & ((1u << PBITS) - 1u)] = (uint16_t)(E_LONG | (1u << 10));   // long-code marker (len field != 0)
      }
    }
  }
  __syncwarp();
  return true;
}

// Decodes one symbol with the LUT, falling back to the canonical walk for codes longer than PBITS.
// Returns the 16-bit entry (kind | len | value); kind LONG with len 0 means invalid code.
template <int PBITS, bool IS_LITLEN>
__device__ __forceinline__ uint32_t decode_sym(uint32_t bits, const uint16_t* lut, const uint16_t* sorted,
                                               const uint16_t* first, const uint16_t* offs, const uint16_t* cnt) {
  uint32_t e = lut[bits & ((1u << PBITS) - 1u)];
  if (e >= E_LONG) {
    if ((e & (15u << 10)) == 0) return E_LONG;          // invalid
    uint32_t rb = __brev(bits);
    e = E_LONG;
    for (uint32_t len = PBITS + 1; len <= 15; len++) {
      uint32_t d = (rb >> (32 - len)) - first[len];
      if (d < cnt[len]) {
        uint32_t s = sorted[offs[len] + d];
        if (IS_LITLEN) e = s < 256 ? (E_LIT | s) : (s == 256 ? E_EOB : (E_SYM | (s - 257)));
        else e = E_SYM | s;
        e |= len << 10;
        break;
      }
    }
  }
  return e;
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
__device__ __forceinline__ uint32_t warp_crc32(const uint8_t* out, uint32_t n, const uint32_t (*tab)[256], int lane) {
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

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(INF_WARPS * 32, 6)
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

  for (;;) {
    uint32_t bi = 0;
    if (lane == 0) bi = atomicAdd(ticket, 1u);
    bi = __shfl_sync(FULL, bi, 0);
    if (bi >= n_blocks) break;
    const BlockDesc bd = blocks[bi];
    uint8_t* out = infl + bd.uoff;
    const uint32_t isize = bd.isize;
    uint32_t outpos = 0, err = INF_OK;
    BitReader br;
    br.init(comp + bd.cdata_off, bd.cdata_len, lane);
    uint32_t mylit = 0, npend = 0, pend_base = 0;   // parked literals: lane i holds out[pend_base + i]
    bool final_block = false;

    while (!final_block && err == INF_OK) {
      uint32_t hdr = br.take(3, lane);
      final_block = hdr & 1u;
      uint32_t btype = hdr >> 1;
      if (btype == 0) {
        // stored: skip to byte boundary, LEN / NLEN, raw copy
        br.consume((8 - (br.bp & 7)) & 7, lane);
        uint32_t len = br.take(16, lane), nlen = br.take(16, lane);
        if ((len ^ nlen) != 0xffffu || outpos + len > isize) { err = INF_ERR_STORED; break; }
        if ((uint32_t)lane < npend) out[pend_base + lane] = (uint8_t)mylit;   // flush parked literals
        npend = 0;
        const uint8_t* src = br.byte_ptr();
        for (uint32_t i = lane; i < len; i += 32) out[outpos + i] = src[i];
        outpos += len;
        uint32_t used = (uint32_t)(src + len - (comp + bd.cdata_off));
        if (used > bd.cdata_len) { err = INF_ERR_INPUT; break; }
        br.init(src + len, bd.cdata_len - used, lane);
        continue;
      }
      if (btype == 3) { err = INF_ERR_BTYPE; break; }
      int n_ll, n_d;
      if (btype == 1) {
        for (int i = lane; i < 288; i += 32) T.cl[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
        if (lane < 30) T.cl[288 + lane] = 5;
        n_ll = 288; n_d = 30;
        __syncwarp();
      } else {
        uint32_t h = br.take(14, lane);
        n_ll = (int)(h & 31u) + 257; n_d = (int)((h >> 5) & 31u) + 1;
        int n_clc = (int)(h >> 10) + 4;
        if (n_ll > 286 || n_d > 30) { err = INF_ERR_TABLE; break; }
        // code-length code: 19 3-bit lengths, staged in cl[] then built into lut_d (7-bit index)
        if (lane < 19) T.cl[lane] = 0;
        __syncwarp();
        for (int i = 0; i < n_clc; i++) {
          uint32_t v = br.take(3, lane);
          if (lane == 0) T.cl[c_clc_order[i]] = (uint8_t)v;
        }
        __syncwarp();
        if (!build_table<7, false>(T.cl, 19, T.lut_d, T.sorted_d, T.first_d, T.offs_d, T.cnt_d, T.nxt, lane)) { err = INF_ERR_TABLE; break; }
        int total = n_ll + n_d, i = 0;
        uint32_t prev = 0;
        uint8_t* cl2 = T.cl + 0;   // decoded lengths overwrite the (no longer needed) pre-code lengths
        while (i < total) {
          uint32_t e = T.lut_d[br.peek() & 127u];
          uint32_t L = (e >> 10) & 15u, s = e & 1023u;
          if (e >= E_LONG || L == 0) { err = INF_ERR_TABLE; break; }
          br.consume(L, lane);
          uint32_t rep, val;
          if (s < 16) { rep = 1; val = s; prev = s; }
          else if (s == 16) { if (i == 0) { err = INF_ERR_TABLE; break; } rep = 3 + br.take(2, lane); val = prev; }
          else if (s == 17) { rep = 3 + br.take(3, lane); val = 0; prev = 0; }
          else { rep = 11 + br.take(7, lane); val = 0; prev = 0; }
          if (i + (int)rep > total) { err = INF_ERR_TABLE; break; }
          for (uint32_t k = lane; k < rep; k += 32) cl2[i + k] = (uint8_t)val;
          i += (int)rep;
        }
        if (err) break;
        __syncwarp();
        if (T.cl[256] == 0) { err = INF_ERR_TABLE; break; }
      }
      if (!build_table<INF_LL_BITS, true>(T.cl, n_ll, T.lut_ll, T.sorted_ll, T.first_ll, T.offs_ll, T.cnt_ll, T.nxt, lane)) { err = INF_ERR_TABLE; break; }
      if (!build_table<INF_D_BITS, false>(T.cl + n_ll, n_d, T.lut_d, T.sorted_d, T.first_d, T.offs_d, T.cnt_d, T.nxt, lane)) { err = INF_ERR_TABLE; break; }

      // ---------------- symbol loop (warp-uniform) ----------------
      for (;;) {
        uint32_t bits = br.peek();
        uint32_t e = decode_sym<INF_LL_BITS, true>(bits, T.lut_ll, T.sorted_ll, T.first_ll, T.offs_ll, T.cnt_ll);
        uint32_t L = (e >> 10) & 15u;
        if (e < E_SYM) {                       // literal
          br.consume(L, lane);
          if (npend == 0) pend_base = outpos;
          if ((uint32_t)lane == npend) mylit = e;
          npend++; outpos++;
          if (npend == 32) {
            if (outpos > isize) { err = INF_ERR_OVERRUN; break; }
            out[pend_base + lane] = (uint8_t)mylit;
            npend = 0;
          }
          continue;
        }
        if (e >= E_LONG) { err = INF_ERR_SYMBOL; break; }
        br.consume(L, lane);
        if (e >= E_EOB) break;                  // end of block
        // length symbol
        uint32_t ls = e & 31u;
        if (ls > 28) { err = INF_ERR_SYMBOL; break; }
        uint32_t len = c_len_base[ls];
        uint32_t xb = c_len_extra[ls];
        if (xb) len += br.take(xb, lane);
        uint32_t de = decode_sym<INF_D_BITS, false>(br.peek(), T.lut_d, T.sorted_d, T.first_d, T.offs_d, T.cnt_d);
        if (de >= E_LONG) { err = INF_ERR_DIST; break; }
        br.consume((de >> 10) & 15u, lane);
        uint32_t ds = de & 31u;
        if (ds > 29) { err = INF_ERR_DIST; break; }
        uint32_t dist = c_dist_base[ds];
        uint32_t dxb = c_dist_extra[ds];
        if (dxb) dist += br.take(dxb, lane);
        if (dist > outpos || outpos + len > isize) { err = dist > outpos ? INF_ERR_DIST : INF_ERR_OVERRUN; break; }
        // flush parked literals, then fence so earlier stores are visible to the copy's loads
        if ((uint32_t)lane < npend) out[pend_base + lane] = (uint8_t)mylit;
        npend = 0;
        __syncwarp();
        uint8_t* dst = out + outpos;
        const uint8_t* src = dst - dist;
        if (dist >= 32) {
          for (uint32_t b0 = 0; b0 < len; b0 += 32) {      // uniform trip count; stripes may feed each other
            uint32_t i = b0 + lane;
            if (i < len) dst[i] = src[i];
            __syncwarp();
          }
        } else if (dist == 1) {
          uint8_t v = src[0];
          for (uint32_t i = lane; i < len; i += 32) dst[i] = v;
        } else {
          for (uint32_t i = lane; i < len; i += 32) dst[i] = src[i % dist];
        }
        outpos += len;
      }
      if (br.overrun && err == INF_OK) err = INF_ERR_INPUT;
    }
    if (err == INF_OK && outpos != isize) err = outpos > isize ? INF_ERR_OVERRUN : INF_ERR_ISIZE;
    if ((uint32_t)lane < npend && err == INF_OK) out[pend_base + lane] = (uint8_t)mylit;
    __syncwarp();
    if (err == INF_OK && check_crc) {
      uint32_t crc = warp_crc32(out, isize, sh.crc_tab, lane);
      if (crc != bd.crc) err = INF_ERR_CRC;
    }
    if (lane == 0) {
      status[bi] = err;
      if (err) atomicCAS(err_flag, 0u, (bi << 4) | err | 0x80000000u);
    }
    __syncwarp();
  }
}

}  // namespace bamscan
