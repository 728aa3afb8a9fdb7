// deflate_sim.cpp -- host model of bgzf_deflate_kernel's LZ77 stage (kernels_deflate.cuh, phase A) + exact dynamic-Huffman cost:
// the compression ratio of parser variants is explored here, on the CPU, before a variant is built on the GPU.
//   usage: deflate_sim <inflated-stream-file> [members]     (prints bytes per variant next to zlib level 6 / 1)
// The model follows the kernel step for step: WARPS regions per member, 32 positions per step, a region-local slice of the hash
// table (most recent position per slot) pre-seeded with the bytes in front of the region, the run candidate p - 1, matches
// measured to 32 bytes, lazy evaluation, greedy parse of the step, the step's last token extended to 258 bytes.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <queue>
#include <vector>
#include <zlib.h>

struct Variant { const char* name; int warps; uint32_t htab; uint32_t preseed; bool lazy; int ways; int minmatch; bool lazy2; };

static uint32_t ld4(const uint8_t* b, uint32_t p) { uint32_t v; memcpy(&v, b + p, 4); return v; }
static uint32_t match_len(const uint8_t* b, uint32_t a, uint32_t p, uint32_t maxlen) { uint32_t n = 0; while (n < maxlen && b[a + n] == b[p + n]) n++; return n; }

static void len_symbol(uint32_t L, uint32_t& sym, uint32_t& eb) {
  uint32_t l = L - 3;
  if (L == 258) { sym = 285; eb = 0; }
  else if (l < 8) { sym = 257 + l; eb = 0; }
  else { uint32_t nb = 31 - __builtin_clz(l); eb = nb - 2; sym = 257 + 4 * eb + 4 + ((l >> eb) & 3); }
}
static void dist_symbol(uint32_t D, uint32_t& sym, uint32_t& eb) {
  uint32_t d = D - 1;
  if (d < 4) { sym = d; eb = 0; }
  else { uint32_t nb = 31 - __builtin_clz(d); eb = nb - 1; sym = 2 * nb + ((d >> eb) & 1); }
}

// exact Huffman code lengths (limit 15 by the kernel's procedure), returns sum freq * len
static uint64_t huff_bits(const std::vector<uint32_t>& f) {
  struct N { uint64_t w; int l, r; };
  std::vector<N> nodes; std::vector<int> leaf;
  typedef std::pair<uint64_t, int> P;
  std::priority_queue<P, std::vector<P>, std::greater<P>> q;
  for (size_t i = 0; i < f.size(); i++) if (f[i]) { nodes.push_back({f[i], -1, -1}); q.push({f[i], (int)nodes.size() - 1}); }
  if (nodes.size() == 1) return f.size() ? nodes[0].w : 0;
  while (q.size() > 1) { P a = q.top(); q.pop(); P b = q.top(); q.pop(); nodes.push_back({a.first + b.first, a.second, b.second}); q.push({a.first + b.first, (int)nodes.size() - 1}); }
  std::vector<int> depth(nodes.size(), 0);
  uint64_t bits = 0; std::vector<uint32_t> bl(64, 0); std::vector<std::pair<uint64_t, int>> leaves;
  for (int i = (int)nodes.size() - 1; i >= 0; i--) {
    if (nodes[i].l >= 0) { depth[nodes[i].l] = depth[i] + 1; depth[nodes[i].r] = depth[i] + 1; }
    else { bl[std::min(depth[i], 60)]++; leaves.push_back({nodes[i].w, std::min(depth[i], 60)}); }
  }
  bool over = false; for (int l = 16; l < 64; l++) if (bl[l]) over = true;
  if (!over) { for (auto& lf : leaves) bits += lf.first * lf.second; return bits; }
  for (int l = 16; l < 64; l++) { bl[15] += bl[l]; bl[l] = 0; }
  uint64_t kraft = 0; for (int l = 1; l <= 15; l++) kraft += (uint64_t)bl[l] << (15 - l);
  while (kraft > (1ull << 15)) { bl[15]--; for (int l = 14; l >= 1; l--) if (bl[l]) { bl[l]--; bl[l + 1] += 2; break; } kraft--; }
  std::sort(leaves.begin(), leaves.end());
  size_t i = 0; for (int l = 15; l >= 1; l--) for (uint32_t c = 0; c < bl[l]; c++, i++) bits += leaves[i].first * l;
  return bits;
}

static uint64_t member_bits(const uint8_t* in, uint32_t isize, const Variant& V) {
  std::vector<uint8_t> b(isize + 300, 0); memcpy(b.data(), in, isize);
  std::vector<uint32_t> fl(286, 0), fd(30, 0); uint64_t extra = 0;
  const uint32_t reg_len = (((isize + V.warps - 1) / V.warps) + 31u) & ~31u, slots = V.htab / V.warps;
  for (int w = 0; w < V.warps; w++) {
    const uint32_t rbeg = std::min(isize, (uint32_t)w * reg_len), rend = std::min(isize, rbeg + reg_len);
    std::vector<uint16_t> tab((size_t)slots * V.ways, 0xffff);
    auto slot = [&](uint32_t v) { uint32_t x = V.minmatch == 3 ? (v & 0xffffffu) : v; return (uint32_t)(((uint64_t)(x * 2654435761u) * slots) >> 32); };
    auto insert = [&](uint32_t p) { uint32_t s = slot(ld4(b.data(), p)); for (int k = V.ways - 1; k > 0; k--) tab[s * V.ways + k] = tab[s * V.ways + k - 1]; tab[s * V.ways] = (uint16_t)p; };
    for (uint32_t q = rbeg > V.preseed ? rbeg - V.preseed : 0; q + 4 <= rbeg; q++) insert(q);
    uint32_t skip = 0;
    for (uint32_t pos = rbeg; pos < rend; pos += 32) {
      uint32_t L[32], D[32]; uint16_t cand[32][8];
      for (int l = 0; l < 32; l++) {                                   // all lanes look up before any lane inserts
        uint32_t p = pos + l; L[l] = 1; D[l] = 0;
        for (int k = 0; k < V.ways; k++) cand[l][k] = 0xffff;
        if (p < rend && p + 4 <= rend) { uint32_t s = slot(ld4(b.data(), p)); for (int k = 0; k < V.ways; k++) cand[l][k] = tab[s * V.ways + k]; }
      }
      for (int l = 0; l < 32; l++) { uint32_t p = pos + l; if (p < rend && p + 4 <= rend) insert(p); }
      if (skip >= 32) { skip -= 32; continue; }
      for (int l = 0; l < 32; l++) {
        uint32_t p = pos + l;
        if (!(p < rend && p + 4 <= rend)) continue;
        uint32_t maxlen = std::min(32u, rend - p), best = 0;
        for (int k = 0; k < V.ways; k++) {
          uint32_t c = cand[l][k];
          if (c < p && p - c <= 32768) { uint32_t n = match_len(b.data(), c, p, maxlen); if (n >= (uint32_t)V.minmatch && n > best) { best = n; D[l] = p - c; } }
        }
        if (p > 0 && ld4(b.data(), p - 1) == ld4(b.data(), p) && best < maxlen) { uint32_t n = match_len(b.data(), p - 1, p, maxlen); if (n > best) { best = n; D[l] = 1; } }
        if (best >= (uint32_t)V.minmatch) L[l] = best;
      }
      if (V.lazy) { uint32_t Lc[32]; memcpy(Lc, L, sizeof L); for (int l = 0; l < 31; l++) if (Lc[l] >= (uint32_t)V.minmatch && Lc[l + 1] > Lc[l]) L[l] = 1;
        if (V.lazy2) for (int l = 0; l < 30; l++) if (L[l] >= (uint32_t)V.minmatch && Lc[l + 2] > Lc[l] + 1 && Lc[l + 1] < (uint32_t)V.minmatch) L[l] = 1; }
      int last = -1; uint32_t t = skip;
      std::vector<int> sel;
      while (t < 32 && pos + t < rend) { sel.push_back((int)t); last = (int)t; t += L[t]; }
      if (sel.empty()) { skip = 0; continue; }
      if (L[last] == 32) { uint32_t lp = pos + last, lmax = std::min(258u, rend - lp); L[last] = match_len(b.data(), lp - D[last], lp, lmax); }
      for (int l : sel) {
        if (L[l] >= (uint32_t)V.minmatch && L[l] > 1) { uint32_t sy, eb; len_symbol(L[l], sy, eb); fl[sy]++; extra += eb; dist_symbol(D[l], sy, eb); fd[sy]++; extra += eb; }
        else fl[b[pos + l]]++;
      }
      uint32_t end = (uint32_t)last + L[last];
      skip = end > 32 ? end - 32 : 0;
    }
  }
  fl[256] = 1; if (!fl[0]) fl[0] = 1; if (!fd[0]) fd[0] = 1; if (!fd[1]) fd[1] = 1;
  return 17 + 57 + 316 * 4 + huff_bits(fl) + huff_bits(fd) + extra;
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: deflate_sim stream [members]\n"); return 2; }
  FILE* f = fopen(argv[1], "rb"); if (!f) return 1;
  std::vector<uint8_t> data; { uint8_t buf[1 << 16]; size_t n; while ((n = fread(buf, 1, sizeof buf, f)) > 0) data.insert(data.end(), buf, buf + n); } fclose(f);
  const uint32_t members = std::min<uint32_t>(argc > 2 ? atoi(argv[2]) : 200, (uint32_t)(data.size() / 0xff00));
  const Variant vs[] = {
    {"committed: 12 warps, 16896 entries in 2-way buckets, preseed 1024, lazy", 12, 16896 / 2, 1024, true, 2, 4, false},
    {"one candidate per slot (the kernel before the last change)", 12, 16896, 1024, true, 1, 4, false},
    {"one candidate, no lazy", 12, 16896, 1024, false, 1, 4, false},
    {"2-way buckets (double memory)", 12, 16896, 1024, true, 2, 4, false},
    {"4-way buckets (same memory)", 12, 16896 / 4, 1024, true, 4, 4, false},
    {"one candidate, min match 3 (hash of 3 bytes)", 12, 16896, 1024, true, 1, 3, false},
    {"one candidate, lazy over two positions", 12, 16896, 1024, true, 1, 4, true},
    {"one candidate, 8 warps", 8, 16896, 1024, true, 1, 4, false},
    {"one candidate, 16 warps", 16, 16896, 1024, true, 1, 4, false},
    {"one candidate, preseed 4096", 12, 16896, 4096, true, 1, 4, false},
    {"one candidate, 32768 slots", 12, 32768, 1024, true, 1, 4, false},
    {"3-way buckets (same memory)", 12, 16896 / 3, 1024, true, 3, 4, false},
    {"8-way buckets (same memory)", 12, 16896 / 8, 1024, true, 8, 4, false},
    {"4-way buckets, 8 warps", 8, 16896 / 4, 1024, true, 4, 4, false},
  };
  uint64_t z6 = 0, z1 = 0;
  for (uint32_t m = 0; m < members; m++) {
    for (int lvl : {6, 1}) { z_stream z; memset(&z, 0, sizeof z); deflateInit2(&z, lvl, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY); std::vector<uint8_t> out(80000); z.next_in = data.data() + (size_t)m * 0xff00; z.avail_in = 0xff00; z.next_out = out.data(); z.avail_out = out.size(); deflate(&z, Z_FINISH); (lvl == 6 ? z6 : z1) += z.total_out; deflateEnd(&z); }
  }
  printf("%u members; zlib-6 %llu bytes, zlib-1 %llu (%.3f x)\n", members, (unsigned long long)z6, (unsigned long long)z1, (double)z1 / z6);
  for (const Variant& V : vs) {
    uint64_t bits = 0;
    for (uint32_t m = 0; m < members; m++) bits += (member_bits(data.data() + (size_t)m * 0xff00, 0xff00, V) + 7) / 8 * 8;
    printf("%-72s %10llu bytes  %.4f x zlib-6\n", V.name, (unsigned long long)(bits / 8), (double)(bits / 8) / z6);
  }
  return 0;
}
