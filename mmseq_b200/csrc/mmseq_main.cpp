/* mmseq_main.cpp — the `mmseq` host program: the reference's command line and
 * file formats (src/mmseq.cpp:156-296, :1469-1694 of eturro/mmseq 1.0.11) over
 * the CUDA hot path behind include/mmq.h.
 *
 *   mmseq [OPTIONS...] hits_file output_base
 *
 * Same flags, defaults, validation messages and exit codes as the reference;
 * same output files and column layouts.  Everything numeric runs on the GPU
 * through the C ABI: initial mu and unique hits, EM, the Gibbs sweeps, the
 * prior draws of hit-less isoforms, trace aggregation, Sokal, percentiles and
 * proportion summaries.  The host only parses, formats and writes.
 * Additive options: -gpus INT (shard hit classes over GPUs of this box),
 * -notraces (skip the four *.trace_gibbs.gz text dumps), -batch FILE / -per_gpu INT (many
 * samples dealt to the GPUs, independent chains: BASELINE config 5).
 */
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <charconv>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <limits>
#include <map>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mmq.h"
#include "fmt_g6.h"
#include "huff_gz.h"
#include "trace_writer.h"
#include "hits_loader.h"

#define QUOTE_(x) #x
#define QUOTE(x) QUOTE_(x)
#ifndef VERSION
#define VERSION 1.0.11-b200
#endif

using namespace std;

/* The usage text of the reference (src/mmseq.cpp:156-177) plus the additive options. */
static void printUsage(ostream& out) {
  out << "Usage: mmseq [OPTIONS...] hits_file output_base" << endl
      << endl
      << "Mandatory arguments:" << endl
      << "  hits_file          hits file generated with `bam2hits`\n"
      << "  output_base        base name for output files" << endl
      << endl
      << "Optional arguments:\n"
      << "  -alpha FLOAT       value of alpha in Gamma prior for mu (default: 0.1)" << endl
      << "  -beta FLOAT        value of beta in Gamma prior for mu (default: 0.1)" << endl
      << "  -max_em_iter INT   maximum number of EM iterations (default: 1000)" << endl
      << "  -epsilon FLOAT     minimum loglik ratio between successive EM iterations (default: 0.1)" << endl
      << "  -gibbs_iter INT    number of Gibbs iterations (default: 16384)" << endl
      << "  -gibbs_ss INT      subsampling interval for Gibbs output (default: gibbs_iter/1024)" << endl
      << "  -seed INT          seed for the PRNG in thread 0 (default: 1234)" << endl
      << "  -percentiles STR   comma-separated list of real-scale marginal posterior percentiles to output (default: \"5,25,50,75,95\")" << endl
      << "  -debug             output additional diagnostic files" << endl
      << "  -help              print this help message" << endl
      << "  -version           print the version" << endl
      << "  -gpus INT          (B200 build) shard hit classes over this many GPUs (default: 1)" << endl
      << "  -notraces          (B200 build) do not write the *.trace_gibbs.gz text dumps" << endl
      << "  -batch FILE        (B200 build) many samples: FILE lists one \"hits_file output_base\" pair per line (no positional arguments)" << endl
      << "  -per_gpu INT       (B200 build) batch mode: samples in flight per GPU (default: 4)" << endl
      << endl;
}

[[noreturn]] static void die(const string& msg) {
  cerr << msg << endl;
  exit(1);
}

static void check(int rc, mmq_handle* h, const char* what) {
  if (rc) die(string("Error: ") + what + ": " + mmq_last_error(h));
}

/* "%g" of a double (== operator<< at the stream's default precision 6, what the reference writes).  General case:
 * std::to_chars, which the standard specifies to give printf's characters in the C locale (libstdc++: 2.5x faster than
 * snprintf).  The trace files hold 4e8 of them, so the common case gets a direct path (fmt_g6.h). */
static inline char* fmt_g(char (&buf)[40], double v) { return fmt_g6(buf, v); }

/* gzip text writer with the reference's stream formatting ("%g" == operator<< at precision 6) */
struct GzText {
  gzFile f = nullptr;
  string buf;
  explicit GzText(const string& path) {
    f = gzopen(path.c_str(), "wb");
    if (!f) die("Error: cannot open " + path + " for writing.");
    buf.reserve(1 << 20);
  }
  void put(const string& s) { buf += s; flush_if(); }
  void put(double v) {
    char t[40];
    buf.append(t, (size_t)(fmt_g(t, v) - t));
    flush_if();
  }
  void flush_if() { if (buf.size() > (1 << 20) - 64) flush(); }
  void flush() { if (!buf.empty()) { gzwrite(f, buf.data(), (unsigned)buf.size()); buf.clear(); } }
  ~GzText() { flush(); if (f) gzclose(f); }
};

/* the *.trace_gibbs.gz writer lives in trace_writer.h (shared with the test-support library) */
static void write_trace_gz(const string& path, const vector<string>& ids, const vector<char>& keep, const double* tr, int L) {
  const string err = mmq::write_trace_gz(path, ids, keep, tr, L);
  if (!err.empty()) die(err);
}

static void write_pcts(ostream& ofs, const double* v, size_t np, char last) {
  for (size_t i = 0; i < np; ++i) {
    ofs << v[i];
    ofs << (i == np - 1 ? last : ',');
  }
}

/* Rows of a table formatted by all host threads (each into a string stream with the file stream's
 * formatting state), written in order: the bytes are those of one loop over `ofs`. */
static void write_rows(ostream& ofs, int64_t rows, const function<void(ostream&, int64_t)>& row) {
  const int T = (int)max<int64_t>(1, min<int64_t>(min(32u, max(1u, thread::hardware_concurrency())), rows / 64));
  if (T == 1) { for (int64_t r = 0; r < rows; ++r) row(ofs, r); return; }
  vector<ostringstream> part((size_t)T);
  vector<thread> th;
  for (int t = 0; t < T; ++t) {
    part[(size_t)t].copyfmt(ofs);
    th.emplace_back([&, t] { for (int64_t r = rows * t / T; r < rows * (t + 1) / T; ++r) row(part[(size_t)t], r); });
  }
  for (auto& x : th) x.join();
  for (auto& p : part) { const string blk = p.str(); ofs.write(blk.data(), (streamsize)blk.size()); }
}

struct Shard {
  int device = 0;
  int64_t row0 = 0, row1 = 0;
  mmq_handle* h = nullptr;
};

/* MMQ_TIMING=1: wall-clock of the host program's phases on stderr */
static void phase(const char* name) {
  static const bool on = getenv("MMQ_TIMING") != nullptr;
  static auto t_start = chrono::steady_clock::now();
  static auto t_last = t_start;
  if (!on) return;
  const auto now = chrono::steady_clock::now();
  fprintf(stderr, "[mmseq timing] %-28s %8.3f s  (+%.3f)\n", name, chrono::duration<double>(now - t_start).count(),
          chrono::duration<double>(now - t_last).count());
  t_last = now;
}

/* ---- command line (src/mmseq.cpp:183-296): one table of options, the reference's messages and exit codes ---- */
struct Options {
  double alpha = 0.1, beta = 0.1, epsilon = 0.1;
  int max_em_iter = 1000, gibbs_iter = 16384, trace_length = 1024, gibbs_ss = 16384 / 1024, seed = 1234;
  vector<double> percentiles{5.0, 25.0, 50.0, 75.0, 95.0};
  bool debug = false, notraces = false;
  int ngpus = 1;
  string batch;     /* -batch FILE: one "hits_file output_base" pair per line */
  int per_gpu = 4;  /* -per_gpu INT: samples in flight per GPU in batch mode */
  vector<string> positional;
};

[[noreturn]] static void usage_error(const string& msg) {
  if (!msg.empty()) cerr << msg << "\n";
  printUsage(cerr);
  exit(1);
}

static bool is_power_of_two(int v) { return v > 0 && (v & (v - 1)) == 0; }

static void parse_percentiles(const string& list, vector<double>& out) {
  out.clear();
  size_t pos = 0;
  while (pos <= list.size()) {
    const size_t comma = list.find(',', pos);
    const string item = list.substr(pos, comma == string::npos ? string::npos : comma - pos);
    if (!item.empty()) {
      const double v = strtod(item.c_str(), NULL);
      if (!(v >= 0 && v <= 100)) { cerr << "Percentiles must be in (0,100)\n"; exit(1); }
      out.push_back(v);
    }
    if (comma == string::npos) break;
    pos = comma + 1;
  }
}

static Options parse_command_line(int argc, char** argv) {
  Options o;
  enum Kind { REAL, INT, TEXT, FLAG };
  struct Spec { const char* name; Kind kind; void* target; };
  string pct;
  const Spec specs[] = {
      {"-alpha", REAL, &o.alpha},         {"-beta", REAL, &o.beta},           {"-max_em_iter", INT, &o.max_em_iter},
      {"-epsilon", REAL, &o.epsilon},     {"-gibbs_iter", INT, &o.gibbs_iter}, {"-gibbs_ss", INT, &o.gibbs_ss},
      {"-seed", INT, &o.seed},            {"-percentiles", TEXT, &pct},       {"-gpus", INT, &o.ngpus},
      {"-batch", TEXT, &o.batch},         {"-per_gpu", INT, &o.per_gpu},      {"-notraces", FLAG, &o.notraces},
      {"-debug", FLAG, &o.debug},
  };
  /* options come first; what is left must be the mandatory arguments (none in batch mode) */
  int i = 1;
  for (; i < argc; ++i) {
    const string arg = argv[i];
    if (arg == "-h" || arg == "--help" || arg == "-help") { cerr << "Calculate mmseq expression estimates.\n"; usage_error(""); }
    if (arg == "-v" || arg == "--version" || arg == "-version") { cerr << "mmseq-" << QUOTE(VERSION) << endl; exit(1); }
    const Spec* hit = nullptr;
    for (const Spec& sp : specs) if (arg == sp.name) { hit = &sp; break; }
    if (!hit) break;
    if (hit->kind == FLAG) { *static_cast<bool*>(hit->target) = true; continue; }
    if (i + 1 >= argc) usage_error("Error: option " + arg + " needs a value.");
    const char* val = argv[++i];
    if (hit->kind == REAL) *static_cast<double*>(hit->target) = strtod(val, NULL);
    else if (hit->kind == INT) *static_cast<int*>(hit->target) = atoi(val);
    else *static_cast<string*>(hit->target) = val;
    if (hit->target == &pct) parse_percentiles(pct, o.percentiles);
  }
  for (; i < argc; ++i) o.positional.push_back(argv[i]);
  const size_t want = o.batch.empty() ? 2 : 0;
  if (o.positional.size() != want) {
    if (!o.positional.empty() && o.positional[0][0] == '-') usage_error("Error: unrecognised option " + o.positional[0] + ".");
    usage_error("Error: mandatory arguments missing.");
  }
  /* src/mmseq.cpp:278-296 (a zero -gibbs_ss would divide by zero there; rejected here) */
  if (o.gibbs_ss == 0 || o.gibbs_iter % o.gibbs_ss != 0) usage_error("Error: gibbs_iter must be divisible by gibbs_ss.");
  o.gibbs_ss = o.gibbs_iter / o.trace_length; /* forced, as the reference does (:284) */
  if (o.gibbs_iter <= 0 || o.trace_length <= 0)
    usage_error("Error: no. of iteratons or trace length <= 0. Possible integer overflow - is gibbs_iter too high?");
  if (!is_power_of_two(o.trace_length)) usage_error("Error: gibbs_iter/gibbs_ss must be a power of 2.");
  if (o.gibbs_ss <= 0) usage_error("Error: gibbs_iter must be at least " + to_string(o.trace_length) + " (the trace length).");
  if (o.ngpus < 1) die("Error: -gpus must be >= 1.");
  if (o.per_gpu < 1) die("Error: -per_gpu must be >= 1.");
  return o;
}

/* One sample: hits file in, the nine output files out (src/mmseq.cpp:312-1694).  first_device: the GPU of a
 * single-GPU run, or the first of o.ngpus consecutive devices.  Messages go to out / err (the terminal for a
 * single sample, the sample's log in batch mode). */
static void run_sample(const Options& o, const string& hits_file, const string& output_base, int first_device, ostream& out, ostream& err) {
  const double alpha = o.alpha, beta = o.beta, epsilon = o.epsilon;
  const int max_em_iter = o.max_em_iter, gibbs_iter = o.gibbs_iter, trace_length = o.trace_length, gibbs_ss = o.gibbs_ss, seed = o.seed;
  const vector<double>& percentiles = o.percentiles;
  const bool debug = o.debug, notraces = o.notraces;
  const int ngpus = o.ngpus;

  /* the CUDA contexts come up (1-2 s) while the hits file is being parsed */
  thread warmup([ngpus, first_device] { for (int g = 0; g < ngpus; ++g) mmq_warmup(first_device + g); });

  /* ---- load: header + records -> hit classes (src/hitsio.cpp, src/mmseq.cpp:312-441) */
  mmq::HitsHeader hdr;
  mmq::HitClasses cls;
  {
    string err;
    if (mmq::load_hits_file(hits_file, mmq::LAYOUT_COLLAPSED, hdr, cls, err)) { warmup.join(); die(err); }
  }
  warmup.join();
  out << "Running mmseq with parameters:\n"
       << "  alpha:         " << alpha << endl
       << "  beta:          " << beta << endl
       << "  max_em_iter:   " << max_em_iter << endl
       << "  epsilon:       " << epsilon << endl
       << "  gibbs_iter:    " << gibbs_iter << endl
       << "  gibbs_ss:      " << gibbs_ss << endl
       << "  seed[0]:       " << seed << endl
       << "  debug:         " << debug << endl
       << "  threads:       " << ngpus << " GPU(s)" << endl;
  const int64_t T = (int64_t)hdr.names.size();
  const int64_t n = cls.n, m = cls.m;
  const int64_t N = cls.N;
  const int64_t G = (int64_t)hdr.gene_names.size(), I = (int64_t)hdr.identical.size();
  const int L = trace_length;
  phase("hits file loaded");
  out << "Found " << n << " transcripts in " << m << " transcript combinations.\r" << endl;
  if (n == 0 || m == 0) die("Error: no mapped fragments in the hits file.");

  vector<double> l;
  {
    string err;
    if (mmq::scaled_lengths(hdr, cls, l, err)) die(err);
  }

  /* ---- shards.  The library orders the classes of a shard itself (class plan, mmq_cls.cu), so the
   * host hands them over as they are: one GPU takes the loader's arrays in place; several GPUs take
   * every ngpus-th class (first-appearance order is unrelated to the cost of a class, so the deal is
   * an equal mix).  The Philox counter of a class is its first-appearance index either way
   * (mmq_problem.class_id): the chain does not depend on the number of GPUs. */
  vector<Shard> shards((size_t)ngpus);
  vector<vector<int64_t>> shard_rp((size_t)ngpus), shard_id((size_t)ngpus);
  vector<vector<int32_t>> shard_col((size_t)ngpus), shard_k((size_t)ngpus);
  for (int g = 0; g < ngpus; ++g) {
    Shard& S = shards[(size_t)g];
    S.device = first_device + g;
    mmq_problem p;
    memset(&p, 0, sizeof p);
    p.n = n;
    if (ngpus == 1) {
      p.m = m; p.nnz = cls.row_ptr[(size_t)m];
      p.row_ptr = cls.row_ptr.data();
      p.col = cls.col.data();
      p.k = cls.k.data();
      p.class_id = nullptr; /* class i is row i */
    } else {
      auto& rp = shard_rp[(size_t)g];
      auto& id = shard_id[(size_t)g];
      auto& cc = shard_col[(size_t)g];
      auto& kk = shard_k[(size_t)g];
      rp.push_back(0);
      for (int64_t i = g; i < m; i += ngpus) {
        id.push_back(i);
        kk.push_back(cls.k[(size_t)i]);
        for (int64_t q = cls.row_ptr[(size_t)i]; q < cls.row_ptr[(size_t)i + 1]; ++q) cc.push_back(cls.col[(size_t)q]);
        rp.push_back((int64_t)cc.size());
      }
      p.m = (int64_t)id.size(); p.nnz = rp.back();
      p.row_ptr = rp.data();
      p.col = cc.data();
      p.k = kk.data();
      p.class_id = id.data();
    }
    S.row0 = 0; S.row1 = p.m;
    p.weight = nullptr;
    p.len = l.data();
    p.alpha = alpha; p.beta = beta; p.class_id_base = 0;
    int rc = mmq_create(&p, S.device, &S.h);
    if (rc) die(string("Error: mmq_create: ") + mmq_last_error(nullptr));
  }
  auto on_all = [&](auto fn) { /* run fn(shard index) on every shard concurrently (collectives need all ranks) */
    if (ngpus == 1) { fn(0); return; }
    vector<thread> th;
    for (int g = 0; g < ngpus; ++g) th.emplace_back([&, g] { fn(g); });
    for (auto& t : th) t.join();
  };
  if (ngpus > 1) {
    char uid[128];
    if (mmq_comm_id(uid)) die(string("Error: mmq_comm_id: ") + mmq_last_error(nullptr));
    on_all([&](int g) { check(mmq_comm_init(shards[(size_t)g].h, uid, g, ngpus), shards[(size_t)g].h, "mmq_comm_init"); });
  }
  if (ngpus > 1) { /* Gibbs: all-reduce + Gamma update fused over NVLink peer memory (no NCCL call per sweep) */
    vector<mmq_handle*> hs;
    for (auto& S : shards) hs.push_back(S.h);
    if (mmq_p2p_attach_local(hs.data(), ngpus)) err << "Note: peer access unavailable (" << mmq_last_error(hs[0]) << "); using NCCL for the count exchange." << endl;
  }
  mmq_handle* H0 = shards[0].h;

  /* ---- initial mu, unique hits (src/mmseq.cpp:610-638) */
  vector<int32_t> unique_hits((size_t)n);
  on_all([&](int g) { check(mmq_init_mu(shards[(size_t)g].h, g == 0 ? unique_hits.data() : nullptr), shards[(size_t)g].h, "mmq_init_mu"); });

  /* ---- unique hits of identical sets and genes (src/mmseq.cpp:643-680, src/uh.cpp) */
  vector<int32_t> identical_unique_hits((size_t)I, 0), gene_unique_hits((size_t)G, 0);
  {
    phase("shards on the GPUs");
    err << "Counting unique hits to sets of identical transcripts...";
    vector<int32_t> set_of((size_t)n, -1);
    for (int64_t s = 0; s < I; ++s)
      for (int32_t hidx : hdr.identical[(size_t)s]) { int32_t c = cls.hdr2col[(size_t)hidx]; if (c >= 0) set_of[(size_t)c] = (int32_t)s; }
    if (I > 0) on_all([&](int g) { check(mmq_unique_hits_sets(shards[(size_t)g].h, set_of.data(), I, g == 0 ? identical_unique_hits.data() : vector<int32_t>((size_t)I).data()), shards[(size_t)g].h, "mmq_unique_hits_sets"); });
    err << "done." << endl;
    err << "Counting unique hits to genes...";
    for (int64_t t = 0; t < n; ++t) set_of[(size_t)t] = hdr.gene_of[(size_t)cls.col2hdr[(size_t)t]];
    on_all([&](int g) { check(mmq_unique_hits_sets(shards[(size_t)g].h, set_of.data(), G, g == 0 ? gene_unique_hits.data() : vector<int32_t>((size_t)G).data()), shards[(size_t)g].h, "mmq_unique_hits_sets"); });
    err << "done." << endl;
  }

  /* ---- .k and .M (src/mmseq.cpp:682-695): one integer per line / "row<TAB>col" per nonzero.  The
   * reference streams them through operator<< with endl; the same bytes are formatted here by all
   * host threads into blocks that are written in order. */
  {
    auto put_int = [](string& out, int64_t v) {
      char buf[24];
      int len = 0;
      if (v < 0) { out.push_back('-'); v = -v; }
      do { buf[len++] = (char)('0' + v % 10); v /= 10; } while (v);
      while (len) out.push_back(buf[--len]);
    };
    auto write_blocks = [&](const string& path, const string& head, int64_t rows, const function<void(string&, int64_t)>& row) {
      FILE* f = fopen(path.c_str(), "wb");
      if (!f) die("Error: cannot write " + path);
      if (!head.empty()) fwrite(head.data(), 1, head.size(), f);
      const int nt = (int)max<int64_t>(1, min<int64_t>(min(32u, max(1u, thread::hardware_concurrency())), rows / 65536));
      vector<string> part((size_t)nt);
      vector<thread> th;
      for (int t = 0; t < nt; ++t)
        th.emplace_back([&, t] {
          string& out = part[(size_t)t];
          const int64_t a = rows * t / nt, b = rows * (t + 1) / nt;
          for (int64_t i = a; i < b; ++i) row(out, i);
        });
      for (auto& x : th) x.join();
      for (const string& blk : part) fwrite(blk.data(), 1, blk.size(), f);
      fclose(f);
    };
    write_blocks(output_base + ".k", "", m, [&](string& out, int64_t i) { put_int(out, cls.k[(size_t)i]); out.push_back('\n'); });
    string head = "#";
    for (int64_t t = 0; t < n; t++) { head += "\t"; head += hdr.names[(size_t)cls.col2hdr[(size_t)t]]; }
    head += "\n";
    write_blocks(output_base + ".M", head, m, [&](string& out, int64_t i) {
      for (int64_t q = cls.row_ptr[(size_t)i]; q < cls.row_ptr[(size_t)i + 1]; ++q) {
        put_int(out, i); out.push_back('\t'); put_int(out, cls.col[(size_t)q]); out.push_back('\n');
      }
    });
  }
  ofstream ofs;

  phase(".k and .M written");
  if (debug) { /* src/mmseq.cpp:697-731 */
    vector<vector<int>> counts_shared((size_t)n, vector<int>(100, 0));
    for (int64_t i = 0; i < m; ++i) {
      const int64_t d = cls.row_ptr[(size_t)i + 1] - cls.row_ptr[(size_t)i];
      for (int64_t q = cls.row_ptr[(size_t)i]; q < cls.row_ptr[(size_t)i + 1]; ++q)
        counts_shared[(size_t)cls.col[(size_t)q]][(size_t)min<int64_t>(d, 100) - 1] += cls.k[(size_t)i];
    }
    ofs.open((output_base + ".sharedcounts").c_str());
    for (int64_t hI = 0; hI < T; ++hI) {
      ofs << hdr.names[(size_t)hI] << "\t";
      const int32_t c = cls.hdr2col[(size_t)hI];
      for (int i = 0; i < 100; i++) { if (c >= 0) ofs << counts_shared[(size_t)c][(size_t)i] << "\t"; else ofs << "0\t"; }
      ofs << endl;
    }
    ofs.close(); ofs.clear();
    ofs.open((output_base + ".doublehits").c_str());
    for (int64_t i = 0; i < n; i++) ofs << cls.doublehits[(size_t)i] << endl;
    ofs.close(); ofs.clear();
    /* Mt rows (transcripts) equal to the previous row are listed in .dupIDs, the others dumped */
    vector<vector<int32_t>> Mt((size_t)n);
    for (int64_t i = 0; i < m; ++i)
      for (int64_t q = cls.row_ptr[(size_t)i]; q < cls.row_ptr[(size_t)i + 1]; ++q) Mt[(size_t)cls.col[(size_t)q]].push_back((int32_t)i);
    ofs.open((output_base + ".Mt-nodups").c_str());
    ofstream ofs2((output_base + ".dupIDs").c_str());
    for (int64_t t = 0; t < n; ++t) {
      if (t > 0 && Mt[(size_t)t] == Mt[(size_t)t - 1]) ofs2 << hdr.names[(size_t)cls.col2hdr[(size_t)t]] << endl;
      else for (int32_t i : Mt[(size_t)t]) ofs << t << "\t" << i << endl;
    }
    ofs.close(); ofs.clear();
    ofs2.close();
  }

  /* ---- EM (src/mmseq.cpp:741-820) */
  vector<double> mu_em((size_t)n);
  {
    out.precision(5);
    out.setf(ios::fixed, ios::floatfield);
    auto t0 = chrono::steady_clock::now();
    int iters = 0;
    double loglik = 0, llr = 0;
    if (!debug) {
      vector<int> it_g((size_t)ngpus); vector<double> ll_g((size_t)ngpus), llr_g((size_t)ngpus);
      on_all([&](int g) { check(mmq_em(shards[(size_t)g].h, max_em_iter, epsilon, &it_g[(size_t)g], &ll_g[(size_t)g], &llr_g[(size_t)g]), shards[(size_t)g].h, "mmq_em"); });
      iters = it_g[0]; loglik = ll_g[0]; llr = llr_g[0];
      if (iters > 0) out << "EM iteration " << iters - 1 << ", log likelihood ratio: " << llr << "            \r";
    } else { /* one iteration per call so that every mu can be dumped (.trace_em.gz, :764-768) */
      GzText g(output_base + ".trace_em.gz");
      for (int64_t t = 0; t < n; t++) { g.put(hdr.names[(size_t)cls.col2hdr[(size_t)t]]); g.put(" "); }
      g.put("\n");
      vector<double> mu((size_t)n);
      vector<double> ll_g((size_t)ngpus);
      on_all([&](int gi) { check(mmq_loglik(shards[(size_t)gi].h, &ll_g[(size_t)gi]), shards[(size_t)gi].h, "mmq_loglik"); });
      loglik = ll_g[0];
      llr = epsilon + 1;
      while (iters < max_em_iter && llr > epsilon) {
        out << "EM iteration " << iters << flush;
        check(mmq_get_mu(H0, mu.data()), H0, "mmq_get_mu");
        for (int64_t t = 0; t < n; t++) { g.put(mu[(size_t)t]); g.put(" "); }
        g.put("\n");
        vector<double> l2((size_t)ngpus);
        on_all([&](int gi) { check(mmq_em(shards[(size_t)gi].h, 1, -numeric_limits<double>::infinity(), nullptr, &l2[(size_t)gi], nullptr), shards[(size_t)gi].h, "mmq_em"); });
        llr = l2[0] - loglik;
        loglik = l2[0];
        out << ", log likelihood ratio: " << llr << "            \r";
        iters++;
      }
    }
    out << endl;
    out.unsetf(ios::floatfield);
    out.precision(6);
    const double sec = chrono::duration<double>(chrono::steady_clock::now() - t0).count();
    err << "EM: " << iters << " iterations in " << sec << " s" << endl;
    check(mmq_get_mu(H0, mu_em.data()), H0, "mmq_get_mu");
  }

  /* ---- Gibbs (src/mmseq.cpp:822-918) */
  {
    auto t0 = chrono::steady_clock::now();
    out << "Gibbs iteration " << gibbs_iter - 1 << "       \r";
    on_all([&](int g) {
      mmq_handle* h = shards[(size_t)g].h;
      check(mmq_gibbs(h, (uint32_t)seed, 0, gibbs_iter, gibbs_ss, L, MMQ_GIBBS_DEFAULT), h, "mmq_gibbs");
      check(mmq_synchronize(h), h, "mmq_synchronize");
    });
    out << endl;
    const double sec = chrono::duration<double>(chrono::steady_clock::now() - t0).count();
    err << "Gibbs: " << gibbs_iter << " sweeps in " << sec << " s (" << gibbs_iter / sec << " sweeps/s, " << (double)m * gibbs_iter / sec
         << " hit-class allocations/s)" << endl;
  }

  phase("EM + Gibbs");
  out << "Amalgamating transcripts and calculating summary statistics..." << flush;

  /* ---- prior-simulated traces for isoforms without hits (src/mmseq.cpp:971-978), on the device */
  vector<int64_t> unobs;           /* header indices */
  vector<int64_t> unobs_slot((size_t)T, -1);
  for (int64_t hI = 0; hI < T; ++hI)
    if (cls.hdr2col[(size_t)hI] < 0) { unobs_slot[(size_t)hI] = (int64_t)unobs.size(); unobs.push_back(hI); }
  const int64_t U = (int64_t)unobs.size();
  vector<double> simu((size_t)U * (size_t)L);
  {
    vector<double> rate((size_t)U);
    for (int64_t u = 0; u < U; ++u) rate[(size_t)u] = beta + hdr.efflen[(size_t)unobs[(size_t)u]] * (double)N / 1000000000.0;
    if (U > 0 && mmq_prior_draws(shards[0].device, U, unobs.data(), rate.data(), alpha, (uint32_t)seed, L, simu.data()))
      die(string("Error: mmq_prior_draws: ") + mmq_last_error(nullptr));
  }

  /* ---- groups: identical sets (:938-954) and genes (:959-982) */
  {
    vector<int64_t> ptr{0};
    vector<int32_t> mem;
    for (int64_t s = 0; s < I; ++s) {
      for (int32_t hidx : hdr.identical[(size_t)s]) { int32_t c = cls.hdr2col[(size_t)hidx]; if (c >= 0) mem.push_back(c); }
      ptr.push_back((int64_t)mem.size());
    }
    check(mmq_set_groups(H0, MMQ_GROUP_IDENTICAL, I, ptr.data(), mem.data(), nullptr), H0, "mmq_set_groups");
  }
  vector<char> gene_has_simu((size_t)G, 0);
  {
    vector<int64_t> ptr{0};
    vector<int32_t> mem;
    vector<double> extra((size_t)G * (size_t)L, 0.0);
    for (int64_t g = 0; g < G; ++g) {
      for (int32_t hidx : hdr.gene_members[(size_t)g]) {
        int32_t c = cls.hdr2col[(size_t)hidx];
        if (c >= 0) mem.push_back(c);
        else {
          gene_has_simu[(size_t)g] = 1;
          const double* sv = simu.data() + (size_t)unobs_slot[(size_t)hidx] * (size_t)L;
          for (int i = 0; i < L; ++i) extra[(size_t)g * (size_t)L + (size_t)i] += sv[i];
        }
      }
      ptr.push_back((int64_t)mem.size());
    }
    check(mmq_set_groups(H0, MMQ_GROUP_GENE, G, ptr.data(), mem.data(), extra.data()), H0, "mmq_set_groups");
  }

  /* ---- summaries on the device */
  const size_t NP = percentiles.size();
  vector<int32_t> pidx(NP);
  for (size_t i = 0; i < NP; i++) pidx[i] = static_cast<int>(round(percentiles[i] / 100.0 * (trace_length - 1))); /* :1113 */
  struct Summ { vector<double> mean, var, tau, pct; vector<int32_t> win, status; };
  auto summarize = [&](int which, int64_t rows) {
    Summ S;
    S.mean.resize((size_t)rows); S.var.resize((size_t)rows); S.tau.resize((size_t)rows);
    S.win.resize((size_t)rows); S.status.resize((size_t)rows); S.pct.resize((size_t)rows * NP);
    if (rows > 0)
      check(mmq_summarize(H0, which, S.mean.data(), S.var.data(), S.tau.data(), S.win.data(), S.status.data(), (int)NP, pidx.data(), S.pct.data()), H0, "mmq_summarize");
    return S;
  };
  Summ St = summarize(0, n), Si = summarize(1, I), Sg = summarize(2, G);
  /* sd, mcse, iact (:1308-1363) */
  auto finish = [&](const Summ& S, vector<double>& sd, vector<double>& mcse, vector<double>& iact) {
    const size_t R = S.mean.size();
    sd.resize(R); mcse.resize(R); iact.resize(R);
    for (size_t r = 0; r < R; ++r) {
      if (S.status[r] != 0) { mcse[r] = trace_length; iact[r] = NAN; }
      else { mcse[r] = sqrt(S.tau[r] * S.var[r] / trace_length); iact[r] = S.tau[r]; }
      sd[r] = sqrt(S.var[r]);
    }
  };
  vector<double> sd, mumcse, iact, sd_identical, mumcse_identical, iact_identical, sd_gene, mumcse_gene, iact_gene;
  finish(St, sd, mumcse, iact);
  finish(Si, sd_identical, mumcse_identical, iact_identical);
  finish(Sg, sd_gene, mumcse_gene, iact_gene);
  const vector<double>& meanmu = St.mean;
  const vector<double>& meanmu_identical = Si.mean;
  const vector<double>& meanmu_gene = Sg.mean;

  /* proportions of observed transcripts (:985-1008, :1236-1265) */
  vector<double> meanprop((size_t)n), meanprobitprop((size_t)n), ssprobitprop((size_t)n), sdprobitprop((size_t)n), pct_prop((size_t)n * NP);
  vector<double> prop_trace;
  {
    vector<int32_t> gene_of_col((size_t)n);
    vector<uint8_t> multi((size_t)n);
    for (int64_t t = 0; t < n; ++t) {
      const int32_t g = hdr.gene_of[(size_t)cls.col2hdr[(size_t)t]];
      gene_of_col[(size_t)t] = g;
      multi[(size_t)t] = hdr.gene_members[(size_t)g].size() > 1;
    }
    if (!notraces) prop_trace.resize((size_t)n * (size_t)L);
    check(mmq_prop_summaries(H0, gene_of_col.data(), multi.data(), meanprop.data(), meanprobitprop.data(), ssprobitprop.data(), (int)NP, pidx.data(),
                             pct_prop.data(), notraces ? nullptr : prop_trace.data()), H0, "mmq_prop_summaries");
    for (int64_t t = 0; t < n; t++) /* :1260-1265, same expression order */
      sdprobitprop[(size_t)t] = sqrt((ssprobitprop[(size_t)t] - meanprobitprop[(size_t)t] * meanprobitprop[(size_t)t] / trace_length) / (trace_length - 1.0));
    for (int64_t t = 0; t < n; t++) meanprobitprop[(size_t)t] /= trace_length;
  }

  /* gene traces are needed on the host for the hit-less isoforms' proportions and the trace file */
  vector<double> gene_trace((size_t)G * (size_t)L);
  if (G > 0) check(mmq_get_group_trace(H0, 2, gene_trace.data()), H0, "mmq_get_group_trace");

  /* hit-less isoforms: percentiles of the simulated trace, proportion summaries (:1174-1192, :1267-1305) */
  vector<double> pct_simu((size_t)U * NP), pct_prop_simu((size_t)U * NP), meanprop_simu((size_t)U), meanprobitprop_simu((size_t)U), sdprobitprop_simu((size_t)U);
  {
    vector<double> probit_in((size_t)U * (size_t)L), probit_out;
    vector<double> prop_simu((size_t)U * (size_t)L);
    for (int64_t u = 0; u < U; ++u) {
      const int32_t g = hdr.gene_of[(size_t)unobs[(size_t)u]];
      const double* sv = simu.data() + (size_t)u * (size_t)L;
      vector<double> v(sv, sv + L);
      sort(v.begin(), v.end());
      for (size_t j = 0; j < NP; ++j) pct_simu[(size_t)u * NP + j] = v[(size_t)pidx[j]];
      for (int i = 0; i < L; ++i) prop_simu[(size_t)u * (size_t)L + (size_t)i] = sv[i] / gene_trace[(size_t)g * (size_t)L + (size_t)i];
      copy(prop_simu.begin() + (ptrdiff_t)((size_t)u * (size_t)L), prop_simu.begin() + (ptrdiff_t)((size_t)(u + 1) * (size_t)L), v.begin());
      sort(v.begin(), v.end());
      for (size_t j = 0; j < NP; ++j) pct_prop_simu[(size_t)u * NP + j] = v[(size_t)pidx[j]];
    }
    for (int64_t u = 0; u < U; ++u) {
      const int32_t g = hdr.gene_of[(size_t)unobs[(size_t)u]];
      const bool multi = hdr.gene_members[(size_t)g].size() > 1;
      double mp = 0.0, s1 = 0.0, s2 = 0.0;
      for (int i = 0; i < L; ++i) {
        const double p = prop_simu[(size_t)u * (size_t)L + (size_t)i];
        mp += p;
        double temp;
        if (multi) temp = mmq_host_ndtri(min(max(p, 0.000000001), 0.999999999));
        else temp = numeric_limits<double>::infinity();
        s1 += temp;
        s2 += temp * temp;
      }
      meanprop_simu[(size_t)u] = mp / trace_length;
      sdprobitprop_simu[(size_t)u] = sqrt((s2 - s1 * s1 / trace_length) / (trace_length - 1.0));
      meanprobitprop_simu[(size_t)u] = s1 / trace_length;
    }
  }

  phase("summaries");
  /* ---- trace dumps (:829-831, :911-917, :1033-1108) */
  if (!notraces) {
    vector<double> tr((size_t)n * (size_t)L);
    check(mmq_get_trace(H0, tr.data()), H0, "mmq_get_trace");
    vector<string> ids((size_t)n);
    for (int64_t t = 0; t < n; ++t) ids[(size_t)t] = hdr.names[(size_t)cls.col2hdr[(size_t)t]];
    write_trace_gz(output_base + ".trace_gibbs.gz", ids, {}, tr.data(), L);
    write_trace_gz(output_base + ".prop.trace_gibbs.gz", ids, {}, prop_trace.data(), L);
    vector<double> itr((size_t)I * (size_t)L);
    if (I > 0) check(mmq_get_group_trace(H0, 1, itr.data()), H0, "mmq_get_group_trace");
    vector<string> iids((size_t)I);
    vector<char> keep((size_t)I);
    for (int64_t s = 0; s < I; ++s) {
      string id;
      const auto& mem = hdr.identical[(size_t)s];
      for (size_t j = 0; j < mem.size(); ++j) {
        id += hdr.names[(size_t)mem[j]];
        if (hdr.names[(size_t)mem[j]].compare(hdr.names[(size_t)mem.back()]) != 0) id += "+";
      }
      iids[(size_t)s] = id;
      keep[(size_t)s] = isfinite(log(itr[(size_t)s * (size_t)L])) != 0; /* :1045 */
    }
    write_trace_gz(output_base + ".identical.trace_gibbs.gz", iids, keep, itr.data(), L);
    vector<char> gkeep((size_t)G);
    for (int64_t g = 0; g < G; ++g) gkeep[(size_t)g] = isfinite(log(gene_trace[(size_t)g * (size_t)L])) != 0; /* :1074 */
    write_trace_gz(output_base + ".gene.trace_gibbs.gz", hdr.gene_names, gkeep, gene_trace.data(), L);
  }

  phase("trace files");
  /* ---- closed forms for unobserved features and gene lengths (:1372-1395) */
  const double digalpha = mmq_host_digamma(alpha);
  const double sqrtpolygalpha = sqrt(mmq_host_trigamma(alpha));
  vector<double> gene_lengths((size_t)G, 0.0);
  for (int64_t g = 0; g < G; ++g) {
    if (isfinite(meanmu_gene[(size_t)g]) != 0) {
      double sum = 0;
      for (int32_t hidx : hdr.gene_members[(size_t)g]) {
        const int32_t c = cls.hdr2col[(size_t)hidx];
        const double len = hdr.efflen[(size_t)hidx];
        if (c >= 0) {
          gene_lengths[(size_t)g] += len * exp(meanmu[(size_t)c]);
          sum += exp(meanmu[(size_t)c]);
        } else {
          gene_lengths[(size_t)g] += len * exp((digalpha - log(beta + len * (double)N / 1000000000.0)));
          sum += exp((digalpha - log(beta + len * (double)N / 1000000000.0)));
        }
      }
      gene_lengths[(size_t)g] /= sum;
    }
  }

  /* ---- .mmseq (:1469-1554) */
  ofs.open((output_base + ".mmseq").c_str());
  ofs << "# Mapped fragments: " << N << endl;
  ofs << "feature_id\tlog_mu\tsd\tmcse\tiact\teffective_length\ttrue_length\tunique_hits\tmean_proportion\tmean_probit_proportion\tsd_probit_proportion\tlog_mu_em\tobserved\tntranscripts\t";
  ofs << "percentiles";
  write_pcts(ofs, percentiles.data(), NP, '\t');
  ofs << "percentiles_proportion";
  write_pcts(ofs, percentiles.data(), NP, '\n');
  write_rows(ofs, T, [&](ostream& ofs, int64_t hI) {
    const string& name = hdr.names[(size_t)hI];
    const int32_t c = cls.hdr2col[(size_t)hI];
    const size_t ntr = hdr.gene_members[(size_t)hdr.gene_of[(size_t)hI]].size();
    if (c >= 0) {
      ofs << name << "\t" << meanmu[(size_t)c] << "\t" << sd[(size_t)c] << "\t" << mumcse[(size_t)c] << "\t" << iact[(size_t)c] << "\t"
          << hdr.efflen[(size_t)hI] << "\t" << hdr.truelen[(size_t)hI] << "\t" << unique_hits[(size_t)c] << "\t" << meanprop[(size_t)c] << "\t"
          << meanprobitprop[(size_t)c] << "\t" << sdprobitprop[(size_t)c] << "\t" << log(mu_em[(size_t)c]) << "\t"
          << "1" << "\t" << ntr << "\t";
      write_pcts(ofs, St.pct.data() + (size_t)c * NP, NP, '\t');
      write_pcts(ofs, pct_prop.data() + (size_t)c * NP, NP, '\n');
    } else {
      const size_t u = (size_t)unobs_slot[(size_t)hI];
      ofs << name << "\t" << digalpha - log(beta + hdr.efflen[(size_t)hI] * (double)N / 1000000000.0) << "\t" << sqrtpolygalpha << "\t"
          << "0" << "\t" << "1" << "\t" << hdr.efflen[(size_t)hI] << "\t" << hdr.truelen[(size_t)hI] << "\t" << 0 << "\t" << meanprop_simu[u] << "\t"
          << meanprobitprop_simu[u] << "\t" << sdprobitprop_simu[u] << "\t" << "NA" << "\t" << "0" << "\t" << ntr << "\t";
      write_pcts(ofs, pct_simu.data() + u * NP, NP, '\t');
      write_pcts(ofs, pct_prop_simu.data() + u * NP, NP, '\n');
    }
  });
  ofs.close(); ofs.clear();

  /* ---- .identical.mmseq (:1556-1613) */
  ofs.open((output_base + ".identical.mmseq").c_str());
  ofs << "# Mapped fragments: " << N << endl;
  ofs << "feature_id\tlog_mu\tsd\tmcse\tiact\teffective_length\ttrue_length\tunique_hits\tobserved\tntranscripts\t";
  ofs << "percentiles";
  write_pcts(ofs, percentiles.data(), NP, '\n');
  for (int64_t s = 0; s < I; ++s) {
    const auto& mem = hdr.identical[(size_t)s];
    const string& backname = hdr.names[(size_t)mem.back()];
    const int32_t first = mem.front();
    if (isfinite(meanmu_identical[(size_t)s])) {
      for (size_t j = 0; j < mem.size(); ++j) {
        const string& nm = hdr.names[(size_t)mem[j]];
        ofs << nm;
        if (nm.compare(backname) != 0) ofs << "+";
        else ofs << "\t" << meanmu_identical[(size_t)s] << "\t" << sd_identical[(size_t)s] << "\t" << mumcse_identical[(size_t)s] << "\t"
                 << iact_identical[(size_t)s] << "\t" << hdr.efflen[(size_t)first] << "\t" << hdr.truelen[(size_t)first] << "\t"
                 << identical_unique_hits[(size_t)s] << "\t" << "1" << "\t" << mem.size() << "\t";
      }
      write_pcts(ofs, Si.pct.data() + (size_t)s * NP, NP, '\n');
    } else {
      for (size_t j = 0; j < mem.size(); ++j) {
        const string& nm = hdr.names[(size_t)mem[j]];
        ofs << nm;
        if (nm.compare(backname) != 0) ofs << "+";
        else ofs << "\t" << log(mem.size()) + digalpha - log(beta + hdr.efflen[(size_t)mem[j]] * (double)N / 1000000000.0) << "\t" << sqrtpolygalpha
                 << "\t" << "0" << "\t" << "NA" << "\t" << hdr.efflen[(size_t)first] << "\t" << hdr.truelen[(size_t)first] << "\t" << 0 << "\t"
                 << "0" << "\t" << mem.size() << "\t";
      }
      for (size_t i = 0; i < NP; i++) ofs << "NA" << (i == NP - 1 ? "\n" : ",");
    }
  }
  ofs.close(); ofs.clear();

  /* ---- .gene.mmseq (:1615-1669) */
  ofs.open((output_base + ".gene.mmseq").c_str());
  ofs << "# Mapped fragments: " << N << endl;
  ofs << "feature_id\tlog_mu\tsd\tmcse\tiact\teffective_length\ttrue_length\tunique_hits\tntranscripts\tobserved\t";
  ofs << "percentiles";
  write_pcts(ofs, percentiles.data(), NP, '\n');
  write_rows(ofs, G, [&](ostream& ofs, int64_t g) {
    bool obs = false;
    for (int32_t hidx : hdr.gene_members[(size_t)g]) if (cls.hdr2col[(size_t)hidx] >= 0) { obs = true; break; }
    if (obs) {
      ofs << hdr.gene_names[(size_t)g] << "\t" << meanmu_gene[(size_t)g] << "\t" << sd_gene[(size_t)g] << "\t" << mumcse_gene[(size_t)g] << "\t"
          << iact_gene[(size_t)g] << "\t" << gene_lengths[(size_t)g] << "\t" << "NA" << "\t" << gene_unique_hits[(size_t)g] << "\t"
          << hdr.gene_members[(size_t)g].size() << "\t" << "1" << "\t";
    } else {
      ofs << hdr.gene_names[(size_t)g] << "\t" << meanmu_gene[(size_t)g] << "\t" << sd_gene[(size_t)g] << "\t" << sd_gene[(size_t)g] / sqrt(trace_length)
          << "\t" << 1 << "\t" << gene_lengths[(size_t)g] << "\t" << "NA" << "\t" << "0" << "\t" << hdr.gene_members[(size_t)g].size() << "\t" << "0" << "\t";
    }
    write_pcts(ofs, Sg.pct.data() + (size_t)g * NP, NP, '\n');
  });
  ofs.close(); ofs.clear();

  out << "done." << endl;
  for (auto& S : shards) mmq_destroy(S.h);

  phase("tables");
  out << "Output files: " << endl
       << "  " << output_base << ".mmseq" << endl
       << "  " << output_base << ".identical.mmseq" << endl
       << "  " << output_base << ".gene.mmseq" << endl;
  out << "  " << output_base << ".M" << endl << "  " << output_base << ".k" << endl << endl;
  if (!notraces)
    out << "  " << output_base << ".trace_gibbs.gz" << endl
         << "  " << output_base << ".identical.trace_gibbs.gz" << endl
         << "  " << output_base << ".gene.trace_gibbs.gz" << endl
         << "  " << output_base << ".prop.trace_gibbs.gz" << endl
         << endl;
  if (debug) {
    out << endl
         << "  " << output_base << ".trace_em.gz" << endl
         << "  " << output_base << ".sharedcounts" << endl
         << "  " << output_base << ".Mt-nodups" << endl
         << "  " << output_base << ".doublehits" << endl
         << "  " << output_base << ".dupIDs" << endl;
  }
}

/* -batch FILE (config 5: many samples, independent chains): every line of FILE names a hits file and an output base.
 * Samples are dealt to the GPUs of the box, per_gpu of them in flight per GPU (their parsing, host-side formatting
 * and device work overlap); each runs as a single-GPU sample and writes its messages to <output_base>.log. */
static int run_batch(const Options& o) {
  vector<pair<string, string>> samples;
  {
    ifstream f(o.batch.c_str());
    if (!f) die("Error: cannot open batch file " + o.batch + ".");
    string a, b;
    while (f >> a >> b) samples.emplace_back(a, b);
  }
  if (samples.empty()) die("Error: no samples in " + o.batch + ".");
  Options one = o;
  one.ngpus = 1;
  const int workers = (int)min<size_t>(samples.size(), (size_t)o.ngpus * (size_t)o.per_gpu);
  cout << "Batch of " << samples.size() << " samples on " << o.ngpus << " GPU(s), " << o.per_gpu << " in flight per GPU." << endl;
  const auto t0 = chrono::steady_clock::now();
  atomic<size_t> next{0};
  vector<thread> th;
  for (int w = 0; w < workers; ++w)
    th.emplace_back([&, w] {
      for (size_t i = next.fetch_add(1); i < samples.size(); i = next.fetch_add(1)) {
        ostringstream log;
        const auto s0 = chrono::steady_clock::now();
        run_sample(one, samples[i].first, samples[i].second, w % o.ngpus, log, log);
        log << "wall " << chrono::duration<double>(chrono::steady_clock::now() - s0).count() << " s on GPU " << w % o.ngpus << endl;
        ofstream lf((samples[i].second + ".log").c_str());
        lf << log.str();
      }
    });
  for (auto& t : th) t.join();
  const double wall = chrono::duration<double>(chrono::steady_clock::now() - t0).count();
  cout << "Batch done: " << samples.size() << " samples in " << wall << " s (" << samples.size() / wall << " samples/s, "
       << (double)samples.size() * o.gibbs_iter / wall << " Gibbs sweeps/s over all samples)." << endl;
  return 0;
}

int main(int argc, char** argv) {
  phase("start");
  const Options o = parse_command_line(argc, argv);
  if (!o.batch.empty()) {
    run_batch(o);
  } else {
    run_sample(o, o.positional[0], o.positional[1], 0, cout, cerr);
  }
  /* every output file is closed: leave without unwinding gigabytes of host vectors and the CUDA
   * contexts (1-2 s at 30M fragments); exit status 0 as src/mmseq.cpp:1727 */
  cout.flush();
  cerr.flush();
  fflush(nullptr);
  _exit(0);
}
