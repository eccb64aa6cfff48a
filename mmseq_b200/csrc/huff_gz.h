/* huff_gz.h — one gzip member (RFC 1952) holding ONE dynamic-Huffman deflate block (RFC 1951) with literals only.
 *
 * The trace files (src/mmseq.cpp:1033-1108) are 3.7 GB of "%g" text on the config-2 sample: digits of continuous random
 * values, in which string matching finds next to nothing — all of deflate's gain is in the entropy coding (measured: zlib's
 * Z_HUFFMAN_ONLY compresses this text to 1 / 2.22, level 1 with matching to 1 / 2.04, level 6 to 1 / 2.24).  zlib's
 * Huffman-only path still runs at 85 MB/s per thread; this encoder does the same job — histogram, length-limited canonical
 * code, bit packing four symbols per store — several times faster.  Any inflate reads the result (tests/test_cli_args.py
 * decodes it with zlib); the bytes after decompression are the reference's.
 */
#ifndef MMQ_HUFF_GZ_H
#define MMQ_HUFF_GZ_H

#include <stdint.h>
#include <string.h>
#include <zlib.h> /* crc32 */

#include <algorithm>
#include <vector>

namespace mmq {
namespace hgz {

/* code lengths (<= limit) of a Huffman code for the symbols with freq > 0; at least two symbols must have freq > 0 */
inline void huffman_lengths(const uint64_t* freq, int nsym, int limit, uint8_t* len) {
  struct Node { uint64_t w; int left, right; };
  std::vector<Node> nodes;
  std::vector<int> order;
  memset(len, 0, (size_t)nsym);
  for (int s = 0; s < nsym; ++s)
    if (freq[s]) { nodes.push_back({freq[s], -1 - s, 0}); order.push_back((int)nodes.size() - 1); }
  const int nleaf = (int)nodes.size();
  if (nleaf == 1) { len[-1 - nodes[0].left] = 1; return; }
  /* two-queue construction: leaves sorted by weight, internal nodes are produced in non-decreasing order */
  std::sort(order.begin(), order.end(), [&](int a, int b) { return nodes[a].w < nodes[b].w || (nodes[a].w == nodes[b].w && a < b); });
  std::vector<int> internal;
  size_t li = 0, ii = 0;
  auto take = [&]() {
    if (li < order.size() && (ii >= internal.size() || nodes[order[li]].w <= nodes[internal[ii]].w)) return order[li++];
    return internal[ii++];
  };
  for (int k = 0; k < nleaf - 1; ++k) {
    const int a = take(), b = take();
    nodes.push_back({nodes[a].w + nodes[b].w, a, b});
    internal.push_back((int)nodes.size() - 1);
  }
  /* depths */
  std::vector<int> depth(nodes.size(), 0);
  for (int i = (int)nodes.size() - 1; i >= nleaf; --i) {
    depth[nodes[i].left] = depth[i] + 1;
    depth[nodes[i].right] = depth[i] + 1;
  }
  int count[64] = {0};
  int maxd = 0;
  for (int i = 0; i < nleaf; ++i) { count[std::min(depth[i], 63)]++; maxd = std::max(maxd, depth[i]); }
  if (maxd > limit) {
    /* fold the over-long codes into the limit and repair the Kraft sum (the classic fix-up: one code one level up, one
     * shorter code split in two) */
    for (int d = limit + 1; d < 64; ++d) { count[limit] += count[d]; count[d] = 0; }
    uint64_t total = 0;
    for (int d = limit; d > 0; --d) total += (uint64_t)count[d] << (limit - d);
    while (total != (1ull << limit)) {
      count[limit]--;
      for (int d = limit - 1; d > 0; --d)
        if (count[d]) { count[d]--; count[d + 1] += 2; break; }
      total--;
    }
  }
  /* the rarest symbols get the longest codes */
  size_t pos = 0;
  for (int d = std::min(maxd, limit); d > 0; --d)
    for (int c = 0; c < count[d]; ++c) len[-1 - nodes[order[pos++]].left] = (uint8_t)d;
}

inline void canonical_codes(const uint8_t* len, int nsym, uint16_t* code) {
  int count[16] = {0};
  for (int s = 0; s < nsym; ++s) count[len[s]]++;
  count[0] = 0;
  uint32_t next[16], c = 0;
  for (int l = 1; l <= 15; ++l) { c = (c + (uint32_t)count[l - 1]) << 1; next[l] = c; }
  for (int s = 0; s < nsym; ++s) {
    const int l = len[s];
    if (!l) { code[s] = 0; continue; }
    uint32_t v = next[l]++, r = 0;
    for (int i = 0; i < l; ++i) { r = (r << 1) | (v & 1); v >>= 1; } /* deflate packs codes starting from their most significant bit */
    code[s] = (uint16_t)r;
  }
}

struct BitWriter {
  uint8_t* p;
  uint64_t buf = 0;
  int cnt = 0;
  explicit BitWriter(uint8_t* dst) : p(dst) {}
  inline void put(uint32_t v, int n) { /* n <= 32, cnt < 8 on entry after flush */
    buf |= (uint64_t)v << cnt;
    cnt += n;
  }
  inline void flush() { /* writes whole bytes, keeps cnt < 8 */
    memcpy(p, &buf, 8);
    p += cnt >> 3;
    buf >>= cnt & ~7;
    cnt &= 7;
  }
  inline uint8_t* finish() {
    flush();
    if (cnt) { *p++ = (uint8_t)buf; buf = 0; cnt = 0; }
    return p;
  }
};

/* appends one gzip member with the n bytes of text to out */
inline void gz_member(const char* text, size_t n, std::vector<uint8_t>& out) {
  const uint8_t* in = (const uint8_t*)text;
  uint64_t h[4][256];
  memset(h, 0, sizeof h);
  size_t i = 0;
  for (; i + 4 <= n; i += 4) { h[0][in[i]]++; h[1][in[i + 1]]++; h[2][in[i + 2]]++; h[3][in[i + 3]]++; }
  for (; i < n; ++i) h[0][in[i]]++;
  uint64_t freq[257];
  for (int s = 0; s < 256; ++s) freq[s] = h[0][s] + h[1][s] + h[2][s] + h[3][s];
  freq[256] = 1; /* end of block */
  int used = 0;
  for (int s = 0; s < 257; ++s) used += freq[s] != 0;
  if (used < 2) freq[in && n ? (in[0] == 0 ? 1 : 0) : 0] = 1; /* an empty text: a second symbol so that the code is a real tree */
  uint8_t llen[257];
  uint16_t lcode[257];
  huffman_lengths(freq, 257, 15, llen);
  canonical_codes(llen, 257, lcode);
  /* the code lengths themselves: 257 literal / end-of-block lengths and two distance codes of one bit (a complete code, as
   * zlib always sends; never used), each written as a plain code-length symbol 0..15 */
  uint8_t seq[259];
  memcpy(seq, llen, 257);
  seq[257] = 1;
  seq[258] = 1;
  uint64_t cfreq[19] = {0};
  for (int k = 0; k < 259; ++k) cfreq[seq[k]]++;
  uint8_t clen[19];
  uint16_t ccode[19];
  huffman_lengths(cfreq, 19, 7, clen);
  canonical_codes(clen, 19, ccode);
  uint64_t bits = 0;
  for (int s = 0; s < 257; ++s) bits += freq[s] * llen[s];
  const size_t cap = 10 + 8 + (size_t)(bits >> 3) + 259 * 2 + 64 + 8 + 16;
  const size_t base = out.size();
  out.resize(base + cap);
  uint8_t* o = out.data() + base;
  static const uint8_t hdr[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 3};
  memcpy(o, hdr, 10);
  BitWriter bw(o + 10);
  bw.put(1, 1);      /* BFINAL */
  bw.put(2, 2);      /* BTYPE = dynamic Huffman */
  bw.put(0, 5);      /* HLIT  = 257 */
  bw.put(1, 5);      /* HDIST = 2 */
  bw.put(15, 4);     /* HCLEN = 19 */
  bw.flush();
  static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  for (int k = 0; k < 19; ++k) { bw.put(clen[order[k]], 3); bw.flush(); }
  for (int k = 0; k < 259; ++k) { bw.put(ccode[seq[k]], clen[seq[k]]); bw.flush(); }
  /* the data: four symbols (<= 60 bits) per store */
  uint32_t tab[256];
  for (int s = 0; s < 256; ++s) tab[s] = (uint32_t)lcode[s] | ((uint32_t)llen[s] << 16);
  i = 0;
  for (; i + 4 <= n; i += 4) {
    const uint32_t a = tab[in[i]], b = tab[in[i + 1]], c = tab[in[i + 2]], d = tab[in[i + 3]];
    const int la = (int)(a >> 16), lb = (int)(b >> 16), lc = (int)(c >> 16), ld = (int)(d >> 16);
    uint64_t v = (uint64_t)(a & 0xffff);
    v |= (uint64_t)(b & 0xffff) << la;
    v |= (uint64_t)(c & 0xffff) << (la + lb);
    v |= (uint64_t)(d & 0xffff) << (la + lb + lc);
    const int tot = la + lb + lc + ld; /* <= 60; cnt < 8: fits the 64-bit buffer only up to 56 bits, so in two halves */
    if (tot <= 56) {
      bw.buf |= v << bw.cnt;
      bw.cnt += tot;
      bw.flush();
    } else {
      const int half = la + lb;
      bw.buf |= (v & ((1ull << half) - 1)) << bw.cnt;
      bw.cnt += half;
      bw.flush();
      bw.buf |= (v >> half) << bw.cnt;
      bw.cnt += tot - half;
      bw.flush();
    }
  }
  for (; i < n; ++i) { const uint32_t a = tab[in[i]]; bw.put(a & 0xffff, (int)(a >> 16)); bw.flush(); }
  bw.put(lcode[256], llen[256]);
  uint8_t* e = bw.finish();
  uint32_t crc = (uint32_t)crc32(0L, Z_NULL, 0);
  for (size_t q = 0; q < n; q += (size_t)1 << 30) crc = (uint32_t)crc32(crc, in + q, (uInt)std::min<size_t>((size_t)1 << 30, n - q));
  const uint32_t isize = (uint32_t)(n & 0xffffffffu);
  memcpy(e, &crc, 4);
  memcpy(e + 4, &isize, 4);
  out.resize((size_t)(e + 8 - out.data()));
}

}  // namespace hgz
}  // namespace mmq
#endif
