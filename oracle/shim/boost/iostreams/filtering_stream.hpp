/* oracle/shim — a MINIMAL stand-in for boost::iostreams::filtering_stream as used by
 * eturro/mmseq (src/mmseq.cpp gzip trace writers, src/hitsio.cpp zlib reader/writer).
 * Output: text is buffered; pop() (or destruction) deflates it (gzip or zlib framing, per the
 * filter pushed) into the sink.  Input: push(istream&) slurps the source, inflating it when a
 * zlib_decompressor was pushed first.  Test infrastructure only. */
#pragma once
#include <zlib.h>

#include <cstring>
#include <iostream>
#include <limits>
#include <sstream>
#include <string>
#include <vector>

namespace boost { namespace iostreams {
struct output {};
struct input {};
namespace zlib { const int best_speed = 1; const int default_compression = -1; }
struct gzip_compressor { int level; explicit gzip_compressor(int l = -1) : level(l) {} };
struct zlib_compressor { int level; explicit zlib_compressor(int l = -1) : level(l) {} };
struct zlib_decompressor {};

namespace detail {
inline std::string deflate_all(const std::string& in, int level, bool gzip) {
  z_stream zs; std::memset(&zs, 0, sizeof zs);
  deflateInit2(&zs, level, Z_DEFLATED, 15 + (gzip ? 16 : 0), 8, Z_DEFAULT_STRATEGY);
  std::string out; std::vector<unsigned char> buf(1 << 16);
  zs.next_in = (Bytef*)in.data(); zs.avail_in = (uInt)in.size();
  int rc;
  do { zs.next_out = buf.data(); zs.avail_out = (uInt)buf.size(); rc = deflate(&zs, Z_FINISH); out.append((char*)buf.data(), buf.size() - zs.avail_out); } while (rc != Z_STREAM_END);
  deflateEnd(&zs);
  return out;
}
inline std::string inflate_all(const std::string& in) {
  z_stream zs; std::memset(&zs, 0, sizeof zs);
  inflateInit(&zs);
  std::string out; std::vector<unsigned char> buf(1 << 16);
  zs.next_in = (Bytef*)in.data(); zs.avail_in = (uInt)in.size();
  int rc;
  do { zs.next_out = buf.data(); zs.avail_out = (uInt)buf.size(); rc = inflate(&zs, Z_NO_FLUSH); out.append((char*)buf.data(), buf.size() - zs.avail_out); } while (rc == Z_OK);
  inflateEnd(&zs);
  return out;
}
}  // namespace detail

template <class Mode> class filtering_stream;

template <>
class filtering_stream<output> : public std::ostream {
 public:
  filtering_stream() : std::ostream(&buf_), sink_(0), mode_(0), level_(-1) {}
  ~filtering_stream() { pop(); }
  void push(const gzip_compressor& g) { mode_ = 1; level_ = g.level; }
  void push(const zlib_compressor& z) { mode_ = 2; level_ = z.level; }
  void push(std::ostream& s) { sink_ = &s; buf_.str(""); this->clear(); }
  void pop() {
    if (!sink_) return;
    this->flush();
    const std::string text = buf_.str();
    if (mode_ == 0) sink_->write(text.data(), (std::streamsize)text.size());
    else { const std::string z = detail::deflate_all(text, level_, mode_ == 1); sink_->write(z.data(), (std::streamsize)z.size()); }
    sink_->flush();
    sink_ = 0; buf_.str("");
  }
  void reset() { pop(); mode_ = 0; }
 private:
  std::stringbuf buf_; std::ostream* sink_; int mode_, level_;
};
typedef filtering_stream<output> filtering_ostream;

template <>
class filtering_stream<input> : public std::istream {
 public:
  filtering_stream() : std::istream(&buf_), inflate_(false) {}
  void push(const zlib_decompressor&) { inflate_ = true; }
  void push(std::istream& s) {
    std::ostringstream tmp; tmp << s.rdbuf();
    buf_.str(inflate_ ? detail::inflate_all(tmp.str()) : tmp.str());
    this->clear();
  }
  void reset() { inflate_ = false; buf_.str(""); this->clear(); }
 private:
  std::stringbuf buf_; bool inflate_;
};
typedef filtering_stream<input> filtering_istream;
}}  // namespace boost::iostreams
