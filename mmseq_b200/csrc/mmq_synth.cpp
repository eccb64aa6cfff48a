/* mmq_synth.cpp — seeded synthetic transcriptomes and fragment hit lists in the
 * shapes BASELINE.json names.  Harness code (tests,
 * bench): the reference ships no sample data (no .hits, no BAM), so every
 * workload is generated.  The output is what a hits file carries
 * (src/hitsio.cpp:162-240): transcript metadata, gene -> isoforms, and per
 * fragment the list of header transcript indices it maps to.
 *
 * Recipe: genes with 1+Geom(0.3) isoforms (cap 30); effective length
 * LogNormal(ln 1500, 0.8) clipped to [50, 30000], true length = eff + 180;
 * true mu LogNormal(0, 2) with 30 % of transcripts set to 0; fragment origin
 * t with probability proportional to mu_t*len_t; hit set = {t} + each sibling
 * isoform w.p. 0.6 + (w.p. 0.02) one transcript of a random other gene.
 * haplo = 1: every transcript exists as copies _A and _B (T = 2 x base); a
 * fragment's hit on a transcript also hits the other copy w.p. 0.9.
 */
#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <string>
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct SplitMix {
  uint64_t s;
  explicit SplitMix(uint64_t seed) : s(seed) {}
  uint64_t next() {
    uint64_t z = (s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
  }
  double uniform() { return ((double)(next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
  double normal() {
    double u1 = uniform(), u2 = uniform();
    return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
  }
};

struct Synth {
  int64_t T = 0, G = 0, N = 0;
  std::vector<int32_t> gene_of;      /* [T] */
  std::vector<int64_t> gene_ptr;     /* [G+1] members are contiguous header indices */
  std::vector<double> efflen, mu;    /* [T] */
  std::vector<int32_t> truelen;      /* [T] */
  std::vector<int64_t> frag_ptr;     /* [N+1] */
  std::vector<int32_t> frag_tid;     /* header indices */
  std::vector<float> frag_w;         /* optional per-hit weights */
};

}  // namespace

extern "C" {

/* seed fixes the transcriptome (genes, lengths, true mu); frag_seed the fragments drawn from it,
 * so that several shards / samples of one transcriptome can be generated independently. */
void* mmq_synth_create(uint64_t seed, uint64_t frag_seed, int64_t T_base, int64_t N, int haplo, int with_weights, int threads) {
  Synth* S = new Synth();
  SplitMix rng(seed);
  /* base transcriptome */
  std::vector<int32_t> bgene;
  std::vector<int64_t> bptr{0};
  while ((int64_t)bgene.size() < T_base) {
    int iso = 1;
    while (iso < 30 && rng.uniform() >= 0.3) ++iso;
    if ((int64_t)bgene.size() + iso > T_base) iso = (int)(T_base - (int64_t)bgene.size());
    const int32_t g = (int32_t)(bptr.size() - 1);
    for (int j = 0; j < iso; ++j) bgene.push_back(g);
    bptr.push_back((int64_t)bgene.size());
  }
  const int64_t Gb = (int64_t)bptr.size() - 1;
  std::vector<double> blen((size_t)T_base), bmu((size_t)T_base);
  for (int64_t t = 0; t < T_base; ++t) {
    double len = std::exp(std::log(1500.0) + 0.8 * rng.normal());
    len = std::min(30000.0, std::max(50.0, len));
    blen[(size_t)t] = std::floor(len * 10.0) / 10.0; /* survives the 6-significant-digit header */
    double m = std::exp(2.0 * rng.normal());
    if (rng.uniform() < 0.3) m = 0.0;
    bmu[(size_t)t] = m;
  }
  const int copies = haplo ? 2 : 1;
  S->T = T_base * copies;
  S->G = Gb;
  S->N = N;
  S->gene_of.resize((size_t)S->T);
  S->efflen.resize((size_t)S->T);
  S->truelen.resize((size_t)S->T);
  S->mu.resize((size_t)S->T);
  S->gene_ptr.resize((size_t)Gb + 1);
  /* header order: gene by gene, isoform by isoform, haplotype copies adjacent */
  for (int64_t g = 0; g <= Gb; ++g) S->gene_ptr[(size_t)g] = bptr[(size_t)g] * copies;
  for (int64_t t = 0; t < T_base; ++t)
    for (int c = 0; c < copies; ++c) {
      const size_t h = (size_t)(t * copies + c);
      S->gene_of[h] = bgene[(size_t)t];
      S->efflen[h] = blen[(size_t)t];
      S->truelen[h] = (int32_t)(blen[(size_t)t] + 180.0);
      S->mu[h] = haplo ? bmu[(size_t)t] * (c == 0 ? 0.6 : 0.4) : bmu[(size_t)t];
    }
  /* origin distribution */
  std::vector<double> cdf((size_t)S->T);
  double acc = 0.0;
  for (int64_t t = 0; t < S->T; ++t) { acc += S->mu[(size_t)t] * S->efflen[(size_t)t]; cdf[(size_t)t] = acc; }
  if (!(acc > 0.0)) { for (int64_t t = 0; t < S->T; ++t) cdf[(size_t)t] = (double)(t + 1); acc = (double)S->T; }
  /* fragments, generated in independent chunks so that the result does not depend on the thread count */
  const int64_t CH = 1 << 16;
  const int64_t nch = (N + CH - 1) / CH;
  std::vector<std::vector<int32_t>> ctid((size_t)nch);
  std::vector<std::vector<int32_t>> clen((size_t)nch);
  std::vector<std::vector<float>> cw((size_t)nch);
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#else
  threads = 1;
#endif
#pragma omp parallel for schedule(dynamic) num_threads(threads)
  for (int64_t c = 0; c < nch; ++c) {
    SplitMix r((seed + 0x51ed270b1ull * frag_seed) * 0x9e3779b97f4a7c15ull + 0x1234567ull + (uint64_t)c * 0xd1342543de82ef95ull);
    const int64_t f0 = c * CH, f1 = std::min(N, f0 + CH);
    auto& tid = ctid[(size_t)c];
    auto& ln = clen[(size_t)c];
    auto& ww = cw[(size_t)c];
    tid.reserve((size_t)(f1 - f0) * 4);
    ln.reserve((size_t)(f1 - f0));
    std::vector<int32_t> hits;
    for (int64_t f = f0; f < f1; ++f) {
      const double u = r.uniform() * acc;
      int64_t t = (int64_t)(std::lower_bound(cdf.begin(), cdf.end(), u) - cdf.begin());
      if (t >= S->T) t = S->T - 1;
      hits.clear();
      hits.push_back((int32_t)t);
      const int32_t g = S->gene_of[(size_t)t];
      const int64_t b = S->gene_ptr[(size_t)g], e = S->gene_ptr[(size_t)g + 1];
      if (haplo) {
        /* siblings are base isoforms; each chosen base isoform hits copy A, B or both */
        const int64_t tb = t / 2;
        if (r.uniform() < 0.9) hits.push_back((int32_t)(t ^ 1));
        for (int64_t sb = b / 2; sb < e / 2; ++sb) {
          if (sb == tb) continue;
          if (r.uniform() < 0.6) {
            const int first = r.uniform() < 0.5 ? 0 : 1;
            hits.push_back((int32_t)(sb * 2 + first));
            if (r.uniform() < 0.9) hits.push_back((int32_t)(sb * 2 + (first ^ 1)));
          }
        }
      } else {
        for (int64_t s = b; s < e; ++s) {
          if (s == t) continue;
          if (r.uniform() < 0.6) hits.push_back((int32_t)s);
        }
      }
      if (r.uniform() < 0.02) {
        int64_t o = (int64_t)(r.uniform() * (double)S->T);
        if (o >= S->T) o = S->T - 1;
        if (S->gene_of[(size_t)o] != g) hits.push_back((int32_t)o);
      }
      ln.push_back((int32_t)hits.size());
      for (int32_t hh : hits) {
        tid.push_back(hh);
        if (with_weights) ww.push_back((float)std::exp(0.5 * r.normal()));
      }
    }
  }
  S->frag_ptr.resize((size_t)N + 1);
  int64_t total = 0;
  {
    int64_t f = 0;
    for (int64_t c = 0; c < nch; ++c)
      for (int32_t l : clen[(size_t)c]) { S->frag_ptr[(size_t)f++] = total; total += l; }
    S->frag_ptr[(size_t)N] = total;
  }
  S->frag_tid.resize((size_t)total);
  if (with_weights) S->frag_w.resize((size_t)total);
  {
    int64_t off = 0;
    for (int64_t c = 0; c < nch; ++c) {
      if (!ctid[(size_t)c].empty()) std::memcpy(S->frag_tid.data() + off, ctid[(size_t)c].data(), sizeof(int32_t) * ctid[(size_t)c].size());
      if (with_weights && !cw[(size_t)c].empty()) std::memcpy(S->frag_w.data() + off, cw[(size_t)c].data(), sizeof(float) * cw[(size_t)c].size());
      off += (int64_t)ctid[(size_t)c].size();
      std::vector<int32_t>().swap(ctid[(size_t)c]);
      std::vector<float>().swap(cw[(size_t)c]);
    }
  }
  return S;
}

/* Fast writers for large hits files (the Python writers in synth.py are for small cases).
 * schema 0 = text (src/hitsio.cpp:162-187), schema 1 = zlib-binary (:189-240), schema 2 = schema 1 plus one fp32 weight
 * per hit after a record's transcript indices (needs a generator made with weights); names as synth.py. */
static void tname(char* buf, int64_t t, int haplo) {
  if (haplo) snprintf(buf, 40, "T%07lld_%c", (long long)(t / 2), "AB"[t % 2]);
  else snprintf(buf, 40, "T%07lld", (long long)t);
}
int mmq_synth_write_hits(void* p, const char* path, int schema, int haplo) {
  Synth* S = (Synth*)p;
  std::string out;
  out.reserve((size_t)64 << 20);
  char nm[48], tmp[128];
  z_stream zs;
  FILE* f = fopen(path, "wb");
  if (!f) return 1;
  std::vector<unsigned char> zbuf((size_t)8 << 20);
  if (schema == 2 && S->frag_w.empty()) { fclose(f); return 3; }
  if (schema >= 1) { memset(&zs, 0, sizeof zs); if (deflateInit(&zs, 1) != Z_OK) { fclose(f); return 2; } }
  auto flush = [&](bool finish) {
    if (schema == 0) { fwrite(out.data(), 1, out.size(), f); out.clear(); return; }
    zs.next_in = (Bytef*)out.data(); zs.avail_in = (uInt)out.size();
    do {
      zs.next_out = zbuf.data(); zs.avail_out = (uInt)zbuf.size();
      deflate(&zs, finish ? Z_FINISH : Z_NO_FLUSH);
      fwrite(zbuf.data(), 1, zbuf.size() - zs.avail_out, f);
    } while (zs.avail_out == 0);
    out.clear();
  };
  auto u32 = [&](uint32_t v) { out.append((const char*)&v, 4); };
  if (schema == 0) {
    for (int64_t t = 0; t < S->T; ++t) { tname(nm, t, haplo); snprintf(tmp, sizeof tmp, "@TranscriptMetaData\t%s\t%g\t%d\n", nm, S->efflen[(size_t)t], S->truelen[(size_t)t]); out += tmp; }
    for (int64_t g = 0; g < S->G; ++g) {
      snprintf(tmp, sizeof tmp, "@GeneIsoforms\tG%07lld", (long long)g); out += tmp;
      for (int64_t t = S->gene_ptr[(size_t)g]; t < S->gene_ptr[(size_t)g + 1]; ++t) { tname(nm, t, haplo); out += "\t"; out += nm; }
      out += "\n";
    }
    for (int64_t r = 0; r < S->N; ++r) {
      snprintf(tmp, sizeof tmp, ">r%lld\n", (long long)r); out += tmp;
      for (int64_t q = S->frag_ptr[(size_t)r]; q < S->frag_ptr[(size_t)r + 1]; ++q) { tname(nm, S->frag_tid[(size_t)q], haplo); out += nm; out += "\n"; }
      if (out.size() > ((size_t)48 << 20)) flush(false);
    }
  } else {
    out += "MMSEQ_HITSFILE\n"; u32((uint32_t)schema); u32((uint32_t)S->T);
    for (int64_t t = 0; t < S->T; ++t) { tname(nm, t, haplo); out += nm; out += "\n"; snprintf(tmp, sizeof tmp, "%g\n", S->efflen[(size_t)t]); out += tmp; u32((uint32_t)S->truelen[(size_t)t]); }
    u32((uint32_t)S->G);
    for (int64_t g = 0; g < S->G; ++g) {
      snprintf(tmp, sizeof tmp, "G%07lld\n", (long long)g); out += tmp;
      u32((uint32_t)(S->gene_ptr[(size_t)g + 1] - S->gene_ptr[(size_t)g]));
      for (int64_t t = S->gene_ptr[(size_t)g]; t < S->gene_ptr[(size_t)g + 1]; ++t) { tname(nm, t, haplo); out += nm; out += "\n"; }
    }
    u32(0); /* no identical sets */
    std::string prev, name;
    for (int64_t r = 0; r < S->N; ++r) {
      snprintf(tmp, sizeof tmp, "r%lld", (long long)r); name = tmp;
      size_t nb = 0, ne = 0, mn = std::min(prev.size(), name.size());
      while (nb < mn && prev[nb] == name[nb]) ++nb;
      while (nb + ne < mn && prev[prev.size() - 1 - ne] == name[name.size() - 1 - ne]) ++ne;
      auto small = [&](size_t v) { if (v < 255) out.push_back((char)v); else { out.push_back((char)255); u32((uint32_t)v); } };
      if (nb == 0 && ne == 0) { out += name; out += "\n"; }
      else { out += "\n"; small(nb); out.append(name, nb, name.size() - nb - ne); out += "\n"; small(ne); }
      prev = name;
      const int64_t b = S->frag_ptr[(size_t)r], e = S->frag_ptr[(size_t)r + 1];
      u32((uint32_t)(e - b));
      for (int64_t q = b; q < e; ++q) u32((uint32_t)S->frag_tid[(size_t)q]);
      if (schema == 2) out.append((const char*)(S->frag_w.data() + b), sizeof(float) * (size_t)(e - b));
      if (out.size() > ((size_t)48 << 20)) flush(false);
    }
  }
  flush(true);
  if (schema >= 1) deflateEnd(&zs);
  fclose(f);
  return 0;
}

void mmq_synth_destroy(void* p) { delete (Synth*)p; }
int64_t mmq_synth_T(void* p) { return ((Synth*)p)->T; }
int64_t mmq_synth_G(void* p) { return ((Synth*)p)->G; }
int64_t mmq_synth_N(void* p) { return ((Synth*)p)->N; }
int64_t mmq_synth_nhits(void* p) { return (int64_t)((Synth*)p)->frag_tid.size(); }
const int32_t* mmq_synth_gene_of(void* p) { return ((Synth*)p)->gene_of.data(); }
const int64_t* mmq_synth_gene_ptr(void* p) { return ((Synth*)p)->gene_ptr.data(); }
const double* mmq_synth_efflen(void* p) { return ((Synth*)p)->efflen.data(); }
const int32_t* mmq_synth_truelen(void* p) { return ((Synth*)p)->truelen.data(); }
const double* mmq_synth_mu(void* p) { return ((Synth*)p)->mu.data(); }
const int64_t* mmq_synth_frag_ptr(void* p) { return ((Synth*)p)->frag_ptr.data(); }
const int32_t* mmq_synth_frag_tid(void* p) { return ((Synth*)p)->frag_tid.data(); }
const float* mmq_synth_frag_w(void* p) { Synth* S = (Synth*)p; return S->frag_w.empty() ? nullptr : S->frag_w.data(); }

} /* extern "C" */
