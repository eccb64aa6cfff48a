/* trace_writer.h — the *.trace_gibbs.gz files of src/mmseq.cpp:829-831, :911-917, :1033-1108: ids each followed by a space,
 * then one line per recorded sweep with the value of every kept feature, "%g" as operator<< writes it, gzip-compressed.
 * 4.2e8 numbers on the config-2 sample: formatting (fmt_g6.h) and compression (huff_gz.h) run on all host threads, a block of
 * lines per thread and round, as independent gzip members written in order; the file writes of a round overlap the next
 * round's formatting.  Shared by the host program and by libmmq_host (test support: tests/test_cli_args.py reads the files back). */
#ifndef MMQ_TRACE_WRITER_H
#define MMQ_TRACE_WRITER_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "fmt_g6.h"
#include "huff_gz.h"

namespace mmq {

/* "%g" text of one trace line range -> one complete gzip member.  A gzip file is a sequence of
 * members (RFC 1952 2.2); zlib's gzread, gzip(1), R's gzfile and Boost's gzip_decompressor all
 * read the concatenation as one stream, so the members can be produced in parallel.  false on a zlib failure. */
inline bool trace_gz_member(const std::string& text, std::vector<unsigned char>& out) {
  /* The text is digits of continuous random values: string matching finds next to nothing in it, the gain is all in the
   * entropy coding (zlib on trace-like text: Z_HUFFMAN_ONLY 1.5x faster than level 1 with matching AND 9 % smaller, ratio
   * 2.22 against 2.04; level 6: 2.24 at an eighth of the speed).  Default: huff_gz.h, the same Huffman-only coding without
   * zlib's per-symbol overhead.  The decompressed bytes are the reference's either way; MMQ_GZIP_LEVEL=6 gives zlib at the
   * reference's settings back. */
  const char* e = getenv("MMQ_GZIP_LEVEL");
  const int level = e ? atoi(e) : 0;
  out.clear();
  if (level <= 0) { hgz::gz_member(text.data(), text.size(), out); return true; }
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  if (deflateInit2(&zs, level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
  out.resize(deflateBound(&zs, (uLong)text.size()) + 64);
  zs.next_in = (Bytef*)text.data();
  zs.avail_in = (uInt)text.size();
  zs.next_out = out.data();
  zs.avail_out = (uInt)out.size();
  const bool ok = deflate(&zs, Z_FINISH) == Z_STREAM_END;
  out.resize(zs.total_out);
  deflateEnd(&zs);
  return ok;
}

/* one (rows x L) trace, feature-major (tr[r * L + i]); keep: empty, or one flag per feature.  Returns an error message or "". */
inline std::string write_trace_gz(const std::string& path, const std::vector<std::string>& ids, const std::vector<char>& keep, const double* tr, int L) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return "Error: cannot open " + path + " for writing.";
  std::vector<size_t> rows;
  for (size_t r = 0; r < ids.size(); ++r)
    if (keep.empty() || keep[r]) rows.push_back(r);
  bool ok = true;
  {
    std::string head;
    for (size_t r : rows) { head += ids[r]; head += ' '; }
    head += '\n';
    std::vector<unsigned char> z;
    ok = trace_gz_member(head, z);
    fwrite(z.data(), 1, z.size(), f);
  }
  const int T = (int)std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  /* lines per block: every thread gets one, at most about 32 MB of text each, at least one line */
  size_t block_bytes = (size_t)32 << 20;
  if (const char* e = getenv("MMQ_TRACE_BLOCK_BYTES")) block_bytes = (size_t)atol(e); /* tests: several rounds on a small trace */
  const int B = (int)std::max<size_t>(1, std::min<size_t>((size_t)(L + T - 1) / (size_t)T, block_bytes / (rows.size() * 12 + 1)));
  std::vector<std::vector<unsigned char>> z[2];
  z[0].resize((size_t)T);
  z[1].resize((size_t)T);
  std::vector<char> good((size_t)T, 1);
  std::thread writer; /* writes the previous round's members while this round is formatted */
  int round = 0;
  for (int base = 0; base < L; base += T * B, ++round) {
    std::vector<std::vector<unsigned char>>& zr = z[round & 1];
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t) {
      const int i0 = base + t * B, i1 = std::min(L, i0 + B);
      zr[(size_t)t].clear();
      if (i0 >= i1) continue;
      th.emplace_back([&, t, i0, i1] {
        /* the trace is feature-major: walk it row by row, appending to the block's lines side by side, so that every cache
         * line of the trace is read once */
        const int nl = i1 - i0;
        std::vector<std::string> line((size_t)nl);
        for (auto& s : line) s.reserve(rows.size() * 12 + 2);
        char tmp[48];
        for (size_t r : rows) {
          const double* v = tr + r * (size_t)L + (size_t)i0;
          for (int j = 0; j < nl; ++j) {
            char* e = fmt_g6(tmp, v[j]);
            *e++ = ' ';
            line[(size_t)j].append(tmp, (size_t)(e - tmp));
          }
        }
        std::string text;
        size_t total = 0;
        for (auto& s : line) total += s.size() + 1;
        text.reserve(total);
        for (auto& s : line) { text += s; text += '\n'; std::string().swap(s); }
        if (!trace_gz_member(text, zr[(size_t)t])) good[(size_t)t] = 0;
      });
    }
    for (auto& x : th) x.join();
    if (writer.joinable()) writer.join();
    writer = std::thread([&zr, f, T] {
      for (int t = 0; t < T; ++t)
        if (!zr[(size_t)t].empty()) fwrite(zr[(size_t)t].data(), 1, zr[(size_t)t].size(), f);
    });
  }
  if (writer.joinable()) writer.join();
  for (char g : good) ok = ok && g;
  if (fclose(f) != 0 || !ok) return "Error: cannot write " + path + ".";
  return "";
}

}  // namespace mmq
#endif
