/* inflate_par.h — a zlib stream inflated on all host threads.
 *
 * Why: the reference writes a .hits file as ONE zlib stream (src/hitsio.cpp:127, best_speed), so reading it is bound by
 * a single inflate() — 4-5 s for the config-2 file, more than EM + 16384 Gibbs sweeps on the GPU (SURVEY.md section 8,
 * row f rank 1).  A deflate stream has no index, but it can still be decoded in parallel:
 *   1. cut the compressed bytes into chunks; in every chunk but the first, FIND a block start: try each bit offset as
 *      the header of a dynamic-Huffman block (BFINAL = 0, BTYPE = 2) and demand what zlib always emits — three COMPLETE
 *      prefix codes (Kraft sum exactly 1), an end-of-block code, lengths that decode without overflow — then decode that
 *      block and check that a valid header follows.  A random offset passes with negligible probability, and a false
 *      hit is caught in step 3 anyway;
 *   2. decode every chunk from its block start with an UNKNOWN 32 KB window: output is 16-bit symbols, a literal byte
 *      or "byte w of the window" (256 + w), which later matches copy around like any other symbol;
 *   3. a chunk stops exactly at the bit where the next one began (if it runs past it the guess was wrong: the caller
 *      falls back to the serial path); windows are resolved front to back (32 KB per chunk), then all chunks are
 *      translated to bytes in parallel; the Adler-32 of the stream is checked from per-chunk sums.
 * (The two-pass idea is that of pugz / rapidgzip; this is an independent implementation of it.)
 *
 * The decoder is a plain table-driven inflate (RFC 1951): 10-bit primary tables with second-level tables for longer
 * codes, stored / fixed / dynamic blocks.  tests/test_loader.py checks it against zlib on streams of every block type.
 */
#ifndef MMQ_INFLATE_PAR_H
#define MMQ_INFLATE_PAR_H

#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

namespace mmq {
namespace ipar {

struct BitReader {
  const uint8_t* p;
  size_t n;      /* bytes */
  uint64_t pos;  /* bit position */
  uint64_t buf = 0;
  int cnt = 0;   /* valid bits in buf */
  size_t next;   /* next byte to load */
  bool over = false;
  BitReader(const uint8_t* data, size_t bytes, uint64_t bitpos) : p(data), n(bytes), pos(bitpos) {
    next = (size_t)(bitpos >> 3);
    refill();
    const int skip = (int)(bitpos & 7);
    buf >>= skip;
    cnt -= skip;
  }
  inline void refill() {
    if (next + 8 <= n) {
      uint64_t w;
      memcpy(&w, p + next, 8);
      buf |= w << cnt;
      const int take = (63 - cnt) >> 3;
      next += (size_t)take;
      cnt += take * 8;
    } else {
      while (cnt <= 56 && next < n) { buf |= (uint64_t)p[next++] << cnt; cnt += 8; }
    }
  }
  inline uint32_t peek(int k) const { return (uint32_t)(buf & ((1ull << k) - 1)); }
  inline void drop(int k) {
    if (k > cnt) { over = true; cnt = 0; buf = 0; return; }
    buf >>= k; cnt -= k; pos += (uint64_t)k;
  }
  inline uint32_t get(int k) {
    if (cnt < k) refill();
    const uint32_t v = peek(k);
    drop(k);
    return v;
  }
  inline void need(int k) { if (cnt < k) refill(); }
};

/* decoding table: entry = symbol << 8 | code length (1..15), or for a primary slot that leads to a second-level table:
 * offset << 8 | 0x80 | extra index bits.  0 = no code. */
struct Huff {
  static constexpr int PB = 10;
  /* primary table + second-level tables: at most 286 symbols with codes longer than PB, each opening at most 2^5 slots */
  uint32_t t[(1 << PB) + 288 * 32];
  int used = 0;
  int maxlen = 0;
  /* returns false unless the lengths form a complete prefix code (what zlib emits) */
  bool build(const uint8_t* len, int nsym) {
    int count[16] = {0};
    for (int i = 0; i < nsym; ++i) count[len[i]]++;
    count[0] = 0;
    int64_t kraft = 0;
    maxlen = 0;
    for (int l = 1; l <= 15; ++l) { kraft += (int64_t)count[l] << (15 - l); if (count[l]) maxlen = l; }
    if (kraft != (1 << 15) || maxlen == 0) return false;
    uint32_t code = 0, first[16];
    for (int l = 1; l <= 15; ++l) { code = (code + (uint32_t)count[l - 1]) << 1; first[l] = code; }
    /* second-level tables: one per distinct PB-bit prefix of the codes longer than PB */
    memset(t, 0, sizeof(uint32_t) << PB);
    used = 1 << PB;
    if (maxlen > PB) {
      /* how many index bits each prefix needs: the longest code under it */
      uint8_t sub[1 << PB];
      memset(sub, 0, sizeof sub);
      uint32_t nxt[16];
      memcpy(nxt, first, sizeof nxt);
      for (int i = 0; i < nsym; ++i) {
        const int l = len[i];
        if (l <= PB) { if (l) nxt[l]++; continue; }
        const uint32_t c = nxt[l]++;
        const uint32_t pre = rev(c >> (l - PB), PB);
        sub[pre] = std::max<uint8_t>(sub[pre], (uint8_t)(l - PB));
      }
      for (uint32_t pre = 0; pre < (1u << PB); ++pre)
        if (sub[pre]) {
          if (used + (1 << sub[pre]) > (int)(sizeof t / sizeof t[0])) return false;
          t[pre] = ((uint32_t)used << 8) | 0x80u | sub[pre];
          memset(t + used, 0, sizeof(uint32_t) << sub[pre]);
          used += 1 << sub[pre];
        }
    }
    for (int i = 0; i < nsym; ++i) {
      const int l = len[i];
      if (!l) continue;
      const uint32_t c = first[l]++;
      const uint32_t r = rev(c, l);
      const uint32_t e = ((uint32_t)i << 8) | (uint32_t)l;
      if (l <= PB) {
        for (uint32_t k = r; k < (1u << PB); k += 1u << l) t[k] = e;
      } else {
        const uint32_t pre = r & ((1u << PB) - 1);
        const uint32_t s = t[pre];
        const int sb = (int)(s & 0x7f);
        const uint32_t off = s >> 8;
        const uint32_t hi = r >> PB;
        for (uint32_t k = hi; k < (1u << sb); k += 1u << (l - PB)) t[off + k] = e;
      }
    }
    return true;
  }
  static inline uint32_t rev(uint32_t c, int l) {
    uint32_t r = 0;
    for (int i = 0; i < l; ++i) { r = (r << 1) | (c & 1); c >>= 1; }
    return r;
  }
  /* symbol or -1; consumes its bits (the reader must hold >= 15 bits or be at the end of the input) */
  inline int decode(BitReader& br) const {
    uint32_t e = t[br.peek(PB)];
    if (e & 0x80u) {
      const int sb = (int)(e & 0x7f);
      e = t[(e >> 8) + ((br.buf >> PB) & ((1u << sb) - 1))];
    }
    if (!e) return -1;
    br.drop((int)(e & 0xff));
    return (int)(e >> 8);
  }
};

/* growable array of 16-bit symbols without value initialisation (a std::vector would zero every page it grows into) */
struct SymBuf {
  uint16_t* p = nullptr;
  size_t n = 0, cap = 0;
  SymBuf() {}
  SymBuf(const SymBuf&) = delete;
  SymBuf& operator=(const SymBuf&) = delete;
  ~SymBuf() { free(p); }
  bool reserve(size_t want) {
    if (want <= cap) return true;
    uint16_t* q = (uint16_t*)realloc(p, want * sizeof(uint16_t));
    if (!q) return false;
    p = q; cap = want;
    return true;
  }
  void release() { free(p); p = nullptr; n = cap = 0; }
  void swap(SymBuf& o) { std::swap(p, o.p); std::swap(n, o.n); std::swap(cap, o.cap); }
  size_t size() const { return n; }
  void clear() { n = 0; }
};

static const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

/* header of a dynamic block at the reader's position (after BFINAL / BTYPE): the two codes, or false */
inline bool read_dynamic_header(BitReader& br, Huff& lit, Huff& dist) {
  static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  br.need(14);
  const int hlit = (int)br.get(5) + 257, hdist = (int)br.get(5) + 1, hclen = (int)br.get(4) + 4;
  if (hlit > 286 || hdist > 30) return false;
  uint8_t cl[19] = {0};
  for (int i = 0; i < hclen; ++i) cl[order[i]] = (uint8_t)br.get(3);
  if (br.over) return false;
  Huff clh;
  if (!clh.build(cl, 19)) return false;
  uint8_t lens[286 + 30];
  int i = 0;
  while (i < hlit + hdist) {
    br.need(15 + 7);
    const int s = clh.decode(br);
    if (s < 0 || br.over) return false;
    if (s < 16) { lens[i++] = (uint8_t)s; continue; }
    int rep;
    uint8_t v = 0;
    if (s == 16) { if (i == 0) return false; v = lens[i - 1]; rep = 3 + (int)br.get(2); }
    else if (s == 17) rep = 3 + (int)br.get(3);
    else rep = 11 + (int)br.get(7);
    if (i + rep > hlit + hdist) return false;
    while (rep--) lens[i++] = v;
  }
  if (br.over || lens[256] == 0) return false;
  if (!lit.build(lens, hlit)) return false;
  if (!dist.build(lens + hlit, hdist)) return false;
  return true;
}

struct FixedTables {
  Huff lit, dist;
  FixedTables() {
    uint8_t l[288];
    for (int i = 0; i < 144; ++i) l[i] = 8;
    for (int i = 144; i < 256; ++i) l[i] = 9;
    for (int i = 256; i < 280; ++i) l[i] = 7;
    for (int i = 280; i < 288; ++i) l[i] = 8;
    lit.build(l, 288);
    uint8_t d[32];
    for (int i = 0; i < 32; ++i) d[i] = 5;
    dist.build(d, 32);
  }
};

/* Decodes blocks from bit position `start` into 16-bit symbols (byte, or 256 + index into the unknown 32 KB window that
 * precedes the chunk; has_window = false: references before the start are an error).  Stops at a block boundary equal to
 * stop_bit (returns 1), after the final block (returns 2; *end_bit = the bit after it), or returns 0 on any error / when the
 * decoding runs past stop_bit.  max_blocks > 0 limits the number of blocks (used by the block finder; returns 3 when hit). */
inline int decode_chunk_impl(const uint8_t* in, size_t n_in, uint64_t start, uint64_t stop_bit, bool has_window, SymBuf& out,
                             uint64_t* end_bit, int max_blocks) {
  static const FixedTables fixed;
  BitReader br(in, n_in, start);
  Huff lit, dist;
  int blocks = 0;
  for (;;) {
    if (br.pos == stop_bit) { *end_bit = br.pos; return 1; }
    if (br.pos > stop_bit) return 0;
    if (max_blocks && blocks == max_blocks) { *end_bit = br.pos; return 3; }
    br.need(3);
    const uint32_t bfinal = br.get(1), btype = br.get(2);
    if (br.over) return 0;
    if (btype == 3) return 0;
    if (btype == 0) {
      br.need(8);
      br.drop((int)((8 - (br.pos & 7)) & 7)); /* to the byte boundary */
      br.need(32);
      const uint32_t len = br.get(16), nlen = br.get(16);
      if (br.over || (len ^ 0xffffu) != nlen) return 0;
      size_t at = (size_t)(br.pos >> 3);
      if (at + len > n_in) return 0;
      if (!out.reserve(out.n + len)) return 0;
      for (uint32_t i = 0; i < len; ++i) out.p[out.n + i] = in[at + i];
      out.n += len;
      br = BitReader(in, n_in, br.pos + (uint64_t)len * 8);
    } else {
      const Huff *L, *D;
      if (btype == 1) { L = &fixed.lit; D = &fixed.dist; }
      else {
        if (!read_dynamic_header(br, lit, dist)) return 0;
        L = &lit; D = &dist;
      }
      size_t o = out.n;
      if (!out.reserve(o + 600)) return 0;
      uint16_t* q = out.p;
      size_t cap = out.cap;
      for (;;) {
        if (o + 600 > cap) { /* room for two literals and one match of 258 */
          if (!out.reserve(std::max<size_t>(cap * 2, o + (1 << 20)))) return 0;
          cap = out.cap;
          q = out.p;
        }
        br.need(48);
        int s = L->decode(br);
        if (s < 256) {
          if (s < 0) return 0;
          q[o++] = (uint16_t)s;
          s = L->decode(br); /* 48 bits hold three codes: a second symbol without another refill */
          if (s < 256) {
            if (s < 0) return 0;
            q[o++] = (uint16_t)s;
            continue;
          }
          br.need(48);
        }
        if (s == 256) break;
        s -= 257;
        if (s >= 29) return 0;
        const int len = LEN_BASE[s] + (int)br.get(LEN_EXTRA[s]);
        br.need(32);
        const int ds = D->decode(br);
        if (ds < 0 || ds >= 30) return 0;
        const int64_t d = DIST_BASE[ds] + (int64_t)br.get(DIST_EXTRA[ds]);
        if (br.over) return 0;
        const int64_t src0 = (int64_t)o - d;
        if (src0 >= 0) {
          if (d >= len) memcpy(q + o, q + src0, (size_t)len * 2);
          else for (int j = 0; j < len; ++j) q[o + j] = q[src0 + j];
        } else {
          if (!has_window || -src0 > 32768) return 0;
          for (int j = 0; j < len; ++j) {
            const int64_t src = src0 + j;
            q[o + j] = src >= 0 ? q[src] : (uint16_t)(256 + 32768 + src);
          }
        }
        o += (size_t)len;
      }
      out.n = o;
      if (br.over) return 0;
    }
    ++blocks;
    if (bfinal) { *end_bit = br.pos; return 2; }
  }
}

inline int decode_chunk(const uint8_t* in, size_t n_in, uint64_t start, uint64_t stop_bit, bool has_window, SymBuf& out,
                        uint64_t* end_bit, int max_blocks = 0) {
  /* the vector's header is updated for every symbol: keep it on this thread's stack, not next to the other chunks' headers
   * (measured: eight threads appending to neighbouring std::vector objects ran no faster than one) */
  SymBuf local;
  local.swap(out);
  uint64_t eb = 0;
  const int rc = decode_chunk_impl(in, n_in, start, stop_bit, has_window, local, &eb, max_blocks);
  *end_bit = eb;
  out.swap(local);
  return rc;
}

/* first bit offset >= from (and < limit) that starts a verified dynamic block, or UINT64_MAX */
inline uint64_t find_block(const uint8_t* in, size_t n_in, uint64_t from, uint64_t limit) {
  Huff lit, dist;
  SymBuf scratch;
  for (uint64_t b = from; b < limit; ++b) {
    const size_t byte = (size_t)(b >> 3);
    if (byte + 4 > n_in) break;
    uint32_t w;
    memcpy(&w, in + byte, 4);
    w >>= (b & 7);
    if ((w & 7u) != 4u) continue;             /* BFINAL = 0, BTYPE = 2 */
    if (((w >> 3) & 31u) > 29u) continue;     /* HLIT <= 29 */
    if (((w >> 8) & 31u) > 29u) continue;     /* HDIST <= 29 */
    {
      BitReader br(in, n_in, b + 3);
      if (!read_dynamic_header(br, lit, dist)) continue;
    }
    /* decode this block and require that two more valid blocks (or the end of the stream) follow */
    scratch.clear();
    uint64_t endb = 0;
    const int rc = decode_chunk(in, n_in, b, UINT64_MAX, true, scratch, &endb, 3);
    if (rc == 3 || rc == 2) return b;
  }
  return UINT64_MAX;
}

/* Inflate a whole zlib stream (2-byte header, deflate data, Adler-32) on `threads` threads.  Returns false when the stream
 * cannot be handled this way (not zlib / preset dictionary / a chunk boundary guess that did not hold / corrupt data): the
 * caller then uses the serial path, which also produces the proper error. */
struct Bytes { /* uninitialised storage: the pages are first touched by the threads that fill them */
  uint8_t* p = nullptr;
  size_t n = 0;
  Bytes() {}
  Bytes(const Bytes&) = delete;
  Bytes& operator=(const Bytes&) = delete;
  ~Bytes() { free(p); }
  bool alloc(size_t bytes) { free(p); p = (uint8_t*)malloc(bytes ? bytes : 1); n = p ? bytes : 0; return p != nullptr; }
  uint8_t* data() { return p; }
  const uint8_t* data() const { return p; }
  size_t size() const { return n; }
};

inline bool inflate_parallel(const uint8_t* in, size_t n_in, int threads, Bytes& out) {
  if (n_in < 8 || (in[0] & 0x0f) != 8 || ((in[0] << 8) | in[1]) % 31 != 0 || (in[1] & 0x20)) return false;
  const uint8_t* d = in + 2;
  const size_t nd = n_in - 2; /* the decoder never reads past the data it needs; the 4 trailer bytes are harmless slack */
  threads = std::max(1, threads);
  const size_t min_chunk = (size_t)1 << 20;
  int nch = (int)std::min<size_t>((size_t)threads * 4, std::max<size_t>(1, nd / min_chunk));
  std::vector<uint64_t> start((size_t)nch + 1, UINT64_MAX);
  start[0] = 0;
  {
    std::atomic<int> nx(1);
    auto work = [&] {
      for (int c; (c = nx.fetch_add(1)) < nch;) {
        const uint64_t from = (uint64_t)(nd / (size_t)nch * (size_t)c) * 8;
        const uint64_t lim = std::min<uint64_t>((uint64_t)(nd / (size_t)nch * (size_t)(c + 1)) * 8, (uint64_t)nd * 8);
        start[(size_t)c] = find_block(d, nd, from, lim);
      }
    };
    std::vector<std::thread> th;
    for (int i = 1; i < threads; ++i) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
  }
  /* chunks without a block start (e.g. inside a long stored block) merge into their predecessor */
  std::vector<uint64_t> st;
  for (int c = 0; c < nch; ++c)
    if (start[(size_t)c] != UINT64_MAX) st.push_back(start[(size_t)c]);
  nch = (int)st.size();
  st.push_back(UINT64_MAX);
  std::vector<SymBuf> sym((size_t)nch);
  std::vector<int> rc((size_t)nch, 0);
  std::vector<uint64_t> endb((size_t)nch, 0);
  {
    std::atomic<int> nx(0);
    auto work = [&] {
      for (int c; (c = nx.fetch_add(1)) < nch;) {
        const size_t span = (size_t)(((c + 1 < nch ? st[(size_t)c + 1] : (uint64_t)nd * 8) - st[(size_t)c]) >> 3);
        if (!sym[(size_t)c].reserve(span * 3 + 65536)) { rc[(size_t)c] = 0; continue; }
        rc[(size_t)c] = decode_chunk(d, nd, st[(size_t)c], st[(size_t)c + 1], c > 0, sym[(size_t)c], &endb[(size_t)c]);
      }
    };
    std::vector<std::thread> th;
    for (int i = 1; i < threads; ++i) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
  }
  for (int c = 0; c < nch; ++c)
    if (rc[(size_t)c] != (c + 1 < nch ? 1 : 2)) return false;
  /* the Adler-32 follows the final block at the next byte boundary */
  const size_t trailer = (size_t)((endb[(size_t)nch - 1] + 7) >> 3);
  if (trailer + 4 > nd) return false;
  const uint32_t want = ((uint32_t)d[trailer] << 24) | ((uint32_t)d[trailer + 1] << 16) | ((uint32_t)d[trailer + 2] << 8) | d[trailer + 3];
  /* windows, front to back: win[c] = the 32 KB in front of chunk c, resolved */
  std::vector<size_t> off((size_t)nch + 1, 0);
  for (int c = 0; c < nch; ++c) off[(size_t)c + 1] = off[(size_t)c] + sym[(size_t)c].size();
  std::vector<std::vector<uint8_t>> win((size_t)nch);
  for (int c = 1; c < nch; ++c) {
    const SymBuf& s = sym[(size_t)c - 1];
    const std::vector<uint8_t>& pw = win[(size_t)c - 1];
    std::vector<uint8_t>& w = win[(size_t)c];
    w.assign(32768, 0);
    const size_t ns = s.size();
    for (size_t i = 0; i < 32768; ++i) {
      /* byte 32768 - 1 - i positions back from the end of chunk c - 1 */
      const int64_t p = (int64_t)ns - 32768 + (int64_t)i;
      if (p >= 0) { const uint16_t v = s.p[(size_t)p]; w[i] = v < 256 ? (uint8_t)v : (pw.empty() ? 0 : pw[v - 256]); }
      else if (!pw.empty()) w[i] = pw[(size_t)(32768 + p)];
    }
  }
  if (!out.alloc(off[(size_t)nch])) return false;
  std::vector<uint32_t> adl((size_t)nch, 1);
  {
    std::atomic<int> nx(0);
    auto work = [&] {
      for (int c; (c = nx.fetch_add(1)) < nch;) {
        SymBuf& s = sym[(size_t)c];
        uint8_t* o = out.data() + off[(size_t)c];
        const uint8_t* w = win[(size_t)c].empty() ? nullptr : win[(size_t)c].data();
        const size_t ns = s.size();
        for (size_t i = 0; i < ns; ++i) { const uint16_t v = s.p[i]; o[i] = v < 256 ? (uint8_t)v : w[v - 256]; }
        s.release();
        uint32_t a = 1;
        for (size_t i = 0; i < ns; i += (size_t)1 << 30) a = (uint32_t)adler32(a, o + i, (uInt)std::min<size_t>((size_t)1 << 30, ns - i));
        adl[(size_t)c] = a;
      }
    };
    std::vector<std::thread> th;
    for (int i = 1; i < threads; ++i) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
  }
  uint32_t a = 1;
  for (int c = 0; c < nch; ++c) a = c == 0 ? adl[0] : (uint32_t)adler32_combine(a, adl[(size_t)c], (z_off_t)(off[(size_t)c + 1] - off[(size_t)c]));
  return a == want;
}

}  // namespace ipar
}  // namespace mmq
#endif
