/* mmseq_oracle.cpp — CPU ORACLE for the mmseq EM + Gibbs hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 * The product (mmseq_b200/csrc, the `mmseq` host program) never links, imports
 * or executes anything under oracle/.
 *
 * PARITY STATUS.  The reference (eturro/mmseq 1.0.11, /root/reference) ships no tests, golden
 * vectors or fixtures for this path (test.sh:1-8 only runs make), and its real dependencies
 * (Boost uBLAS/iostreams, GSL) are absent from this image.  Two pieces of the reference's OWN
 * code are nevertheless compiled from where they lie (oracle/Makefile, outputs in oracle/_ref/):
 *   - src/sokal.cc, self-contained: orc_sokal below and the device kernel are pinned against it
 *     (tests/golden/sokal_golden.npz, tests/test_oracle_sokal.py);
 *   - src/mmseq.cpp + src/hitsio.cpp + src/uh.cpp + src/sokal.cc, UNMODIFIED, against
 *     oracle/shim/ (minimal stand-ins for the few Boost/GSL entry points they use) ->
 *     oracle/_ref/mmseq_ref.  Its outputs for a small hits file are committed under
 *     tests/golden/ref_small/ and pin everything that does not depend on GSL's bit stream:
 *     class construction and numbering (.k, .M byte for byte), unique hits of transcripts / genes /
 *     identical sets, the EM estimate (log_mu_em), closed forms, table layout and formatting
 *     (tests/test_reference_run.py).  The stochastic columns are compared within Monte-Carlo
 *     standard error.
 * What stays UNPINNED: GSL's own arithmetic (sampler bit streams, special-function rounding) —
 * restated from the published algorithms in oracle/gsl_like.h, the same code the shim links —
 * and therefore the reference's exact random stream, which depends on the OpenMP thread count
 * anyway (src/mmseq.cpp:834-838).
 *
 * Third-party arithmetic the reference takes from GSL (unpinned version,
 * src/Makefile:16 -lgsl; .travis.yml:12 libgsl-dev on bionic => 2.4) is
 * restated here from the published algorithms:
 *   gsl_rng_mt19937       Matsumoto & Nishimura 1998 (std::mt19937 is the same
 *                         generator and the same 2002 seeding recurrence)
 *   gsl_ran_multinomial   conditional binomials (Davis 1993), zero categories skipped
 *   gsl_ran_binomial      BTPE (Kachitvichyanukul & Schmeiser 1988), inversion below mean 14
 *   gsl_ran_gamma         Marsaglia & Tsang 2000; a<1 boost U^(1/a)
 *   gsl_ran_gaussian*     GSL uses a ziggurat inside gsl_ran_gamma; the polar
 *                         method is used here (same distribution, other stream)
 * Two Gibbs chains are provided:
 *   orc_gibbs_replay   the shared Philox stream of include/mmq_sampler.h, same
 *                      arithmetic as the kernels => integer outputs bit-exact
 *   orc_gibbs_gsl      the reference's own data flow (src/mmseq.cpp:851-918):
 *                      one MT19937 per OpenMP thread seeded seed+thread, dense
 *                      per-thread count partials summed after a barrier; used
 *                      for statistical comparison and as the CPU baseline.
 */
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>
#ifdef _OPENMP
#include <omp.h>

#include "../mmseq_b200/csrc/mmq_cls_plan.h" /* the product's host-side class plan: built and replayed here for the CPU tests */
#endif

#include "../include/mmq_sampler.h"
#include "gsl_like.h"

extern "C" {

/* ------------------------------------------------------------------------
 * Philox known-answer hook (Random123 kat_vectors: philox4x32 10 rounds). */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
  mmq_philox4x32_10(c, key[0], key[1]);
  for (int i = 0; i < 4; ++i) out[i] = c[i];
}

/* Sampler hooks for distribution tests (shared-header samplers). */
void orc_draw_uniform(uint32_t seed, int64_t cnt, double* out) {
  for (int64_t i = 0; i < cnt; ++i) {
    mmq_rng g;
    mmq_rng_init(&g, seed, MMQ_STREAM_GAMMA, (uint64_t)i, 0);
    out[i] = mmq_uniform(&g);
  }
}
void orc_draw_normal(uint32_t seed, int64_t cnt, double* out) {
  for (int64_t i = 0; i < cnt; ++i) {
    mmq_rng g;
    mmq_rng_init(&g, seed, MMQ_STREAM_GAMMA, (uint64_t)i, 0);
    uint32_t w[4];
    mmq_rng_block(&g, 1u, w);
    out[i] = mmq_normal_bm(w[0], w[1], w[2]);
  }
}
void orc_draw_gamma(uint32_t seed, int64_t cnt, double a, double rate, double* out) {
  for (int64_t i = 0; i < cnt; ++i) {
    mmq_rng g;
    mmq_rng_init(&g, seed, MMQ_STREAM_GAMMA, (uint64_t)i, 0);
    out[i] = mmq_gamma(&g, a, rate);
  }
}
void orc_draw_binomial(uint32_t seed, int64_t cnt, int64_t n, double p, int64_t* out) {
  for (int64_t i = 0; i < cnt; ++i) {
    mmq_rng g;
    mmq_rng_init(&g, seed, MMQ_STREAM_ALLOC, (uint64_t)i, 0);
    out[i] = mmq_binomial(&g, 0u, n, p);
  }
}
void orc_draw_alloc(uint32_t seed, int64_t cnt, int d, const double* p, int64_t k, int32_t* out) {
  for (int64_t i = 0; i < cnt; ++i) mmq_alloc_row(p, out + i * d, d, k, seed, (uint64_t)i, 0);
}
void orc_math(int which, int64_t cnt, const double* in, double* out) {
  for (int64_t i = 0; i < cnt; ++i)
    out[i] = which == 0 ? mmq_log(in[i]) : which == 1 ? mmq_exp(in[i]) : which == 2 ? mmq_ndtri(in[i]) : which == 3 ? mmq_log1p(in[i])
           : mmq_cos2pi_u32((uint32_t)in[i]); /* 4: cos(2 pi (w + 1/2) / 2^32), w passed as a double */
}

/* ------------------------------------------------------------------------
 * Initial mu and the shared-count histogram.  src/mmseq.cpp:610-638:
 *   mu[t] = (sum over classes i containing t of k[i]/|i|) / l[t]
 *   counts_shared[t][min(|i|,100)-1] += k[i]      (bin 0 = unique_hits)
 * The reference walks column t of M top to bottom, i.e. classes in ascending
 * row index; the same summation order is used here. */
void orc_init_mu(int64_t m, int64_t n, const int64_t* rp, const int32_t* col, const int32_t* k,
                 const double* l, double* mu, int32_t* unique_hits, int32_t* counts_shared) {
  for (int64_t t = 0; t < n; ++t) mu[t] = 0.0;
  if (unique_hits) for (int64_t t = 0; t < n; ++t) unique_hits[t] = 0;
  if (counts_shared) std::memset(counts_shared, 0, sizeof(int32_t) * (size_t)n * 100);
  for (int64_t i = 0; i < m; ++i) {
    const int64_t d = rp[i + 1] - rp[i];
    const int32_t ki = k ? k[i] : 1;
    for (int64_t q = rp[i]; q < rp[i + 1]; ++q) {
      const int32_t t = col[q];
      mu[t] += (double)ki / (double)d;
      if (counts_shared) counts_shared[(size_t)t * 100 + (size_t)std::min<int64_t>(d, 100) - 1] += ki;
      if (unique_hits && d == 1) unique_hits[t] += ki;
    }
  }
  for (int64_t t = 0; t < n; ++t) mu[t] /= l[t];
}

/* inner_prod(row i of M, mu): ascending column order (uBLAS sparse row iteration). */
static inline double row_dot(const int64_t* rp, const int32_t* col, const float* w, const double* mu, int64_t i) {
  double s = 0.0;
  if (w) for (int64_t q = rp[i]; q < rp[i + 1]; ++q) s += (double)w[q] * mu[col[q]];
  else for (int64_t q = rp[i]; q < rp[i + 1]; ++q) s += mu[col[q]];
  return s;
}

/* Log-likelihood.  src/mmseq.cpp:745-754 (and :796-802):
 *   sum_i k[i] log(inner_prod(M_i, mu)) - sum_t mu[t] l[t] */
double orc_loglik(int64_t m, int64_t n, const int64_t* rp, const int32_t* col, const int32_t* k,
                  const float* w, const double* l, const double* mu) {
  double ll = 0.0;
  for (int64_t i = 0; i < m; ++i) ll += (double)(k ? k[i] : 1) * std::log(row_dot(rp, col, w, mu, i));
  for (int64_t t = 0; t < n; ++t) ll -= mu[t] * l[t];
  return ll;
}

/* EM.  src/mmseq.cpp:756-811.  llr starts at epsilon+1; loop while
 * iter < max_em_iter && llr > epsilon; update
 *   mu'[t] = mu[t] * (sum_{i containing t} k[i] / inner_prod(M_i, mu)) / l[t]     (:781-794)
 * with classes visited in ascending row index (row t of Mt).  The reference
 * recomputes inner_prod for every (t,i) pair; the value is the same number
 * every time, so it is computed once per class here.  Returns the number of
 * iterations; *loglik_out = final log-likelihood. */
int orc_em(int64_t m, int64_t n, const int64_t* rp, const int32_t* col, const int32_t* k,
           const float* w, const double* l, double* mu, int max_iter, double eps,
           double* loglik_out, double* llr_out) {
  const int64_t nnz = rp[m];
  /* transposed structure: for each t the classes containing it, ascending */
  std::vector<int64_t> tp((size_t)n + 1, 0);
  for (int64_t q = 0; q < nnz; ++q) tp[(size_t)col[q] + 1]++;
  for (int64_t t = 0; t < n; ++t) tp[(size_t)t + 1] += tp[(size_t)t];
  std::vector<int64_t> trow((size_t)nnz), tpos((size_t)nnz), fill(tp.begin(), tp.end() - 1);
  for (int64_t i = 0; i < m; ++i)
    for (int64_t q = rp[i]; q < rp[i + 1]; ++q) {
      int64_t dst = fill[(size_t)col[q]]++;
      trow[(size_t)dst] = i;
      tpos[(size_t)dst] = q;
    }
  std::vector<double> D((size_t)m), mu_temp((size_t)n);
  double loglik = orc_loglik(m, n, rp, col, k, w, l, mu);
  double llr = eps + 1.0;
  int iter = 0;
  while (iter < max_iter && llr > eps) {
    for (int64_t i = 0; i < m; ++i) D[(size_t)i] = row_dot(rp, col, w, mu, i);
    for (int64_t t = 0; t < n; ++t) {
      double sum = 0.0;
      for (int64_t q = tp[(size_t)t]; q < tp[(size_t)t + 1]; ++q) {
        const int64_t i = trow[(size_t)q];
        const double ki = (double)(k ? k[i] : 1);
        sum += (w ? ki * (double)w[tpos[(size_t)q]] : ki) / D[(size_t)i];
      }
      mu_temp[(size_t)t] = mu[t] * sum / l[t];
    }
    const double ll2 = orc_loglik(m, n, rp, col, k, w, l, mu_temp.data());
    for (int64_t t = 0; t < n; ++t) mu[t] = mu_temp[(size_t)t];
    llr = ll2 - loglik;
    loglik = ll2;
    ++iter;
  }
  if (loglik_out) *loglik_out = loglik;
  if (llr_out) *llr_out = llr;
  return iter;
}

/* ------------------------------------------------------------------------
 * Gibbs, shared Philox stream (CPU replay of the kernels).
 * One sweep = src/mmseq.cpp:857-908:
 *   x_i ~ Multinomial(k[i]; mu[cols(i)])              (:865-880)
 *   counts[t] = sum_i x_it                             (:887, :895-899)
 *   mu[t] ~ Gamma(alpha + counts[t], 1/(beta + l[t]))  (:904-908)
 * x (nnz ints, CSR order) and counts are optional outputs of this sweep. */
static void sweep_replay_impl(int64_t m, int64_t n, const int64_t* rp, const int32_t* col, const int32_t* k,
                      const float* w, const double* l, double alpha, double beta, uint32_t seed,
                      uint32_t sweep, int64_t class_id_base, double* mu, int32_t* x_out,
                      int32_t* counts_out, int do_gamma, const int64_t* class_id) {
  std::vector<int32_t> counts((size_t)n, 0);
  std::vector<double> p;
  std::vector<int32_t> x;
  for (int64_t i = 0; i < m; ++i) {
    const int d = (int)(rp[i + 1] - rp[i]);
    p.resize((size_t)d);
    x.resize((size_t)d);
    for (int j = 0; j < d; ++j) {
      const int64_t q = rp[i] + j;
      p[(size_t)j] = w ? mu[col[q]] * (double)w[q] : mu[col[q]];
    }
    mmq_alloc_row(p.data(), x.data(), d, (int64_t)(k ? k[i] : 1), seed, (uint64_t)(class_id ? class_id[i] : class_id_base + i), sweep);
    for (int j = 0; j < d; ++j) {
      const int64_t q = rp[i] + j;
      counts[(size_t)col[q]] += x[(size_t)j];
      if (x_out) x_out[q] = x[(size_t)j];
    }
  }
  if (counts_out) for (int64_t t = 0; t < n; ++t) counts_out[t] = counts[(size_t)t];
  if (do_gamma)
    for (int64_t t = 0; t < n; ++t) {
      mmq_rng g;
      mmq_rng_init(&g, seed, MMQ_STREAM_GAMMA, (uint64_t)t, sweep);
      mu[t] = mmq_gamma(&g, alpha + (double)counts[(size_t)t], beta + l[t]);
    }
}

void orc_sweep_replay(int64_t m, int64_t n, const int64_t* rp, const int32_t* col, const int32_t* k,
                      const float* w, const double* l, double alpha, double beta, uint32_t seed,
                      uint32_t sweep, int64_t class_id_base, double* mu, int32_t* x_out,
                      int32_t* counts_out, int do_gamma) {
  sweep_replay_impl(m, n, rp, col, k, w, l, alpha, beta, seed, sweep, class_id_base, mu, x_out, counts_out, do_gamma, nullptr);
}

/* Same sweep with explicit Philox counters per class (classes handed over in any order). */
void orc_sweep_replay_ids(int64_t m, int64_t n, const int64_t* rp, const int32_t* col, const int32_t* k,
                          const float* w, const double* l, double alpha, double beta, uint32_t seed,
                          uint32_t sweep, const int64_t* class_id, double* mu, int32_t* x_out,
                          int32_t* counts_out, int do_gamma) {
  sweep_replay_impl(m, n, rp, col, k, w, l, alpha, beta, seed, sweep, 0, mu, x_out, counts_out, do_gamma, class_id);
}

/* Allocation step of one sweep on `threads` host threads: counts only (integer sums, so the result does not depend on the
 * thread count).  Same draws as orc_sweep_replay; used by bench.py's correctness gates at full size. */
void orc_sweep_counts(int64_t m, int64_t n, const int64_t* rp, const int32_t* col, const int32_t* k, const float* w,
                      uint32_t seed, uint32_t sweep, int64_t class_id_base, const int64_t* class_id, const double* mu,
                      int32_t* counts_out, int threads) {
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#else
  threads = 1;
#endif
  std::vector<int32_t> part((size_t)n * (size_t)threads, 0);
#pragma omp parallel num_threads(threads)
  {
#ifdef _OPENMP
    int32_t* c = part.data() + (size_t)n * (size_t)omp_get_thread_num();
#else
    int32_t* c = part.data();
#endif
    std::vector<double> p;
    std::vector<int32_t> x;
#pragma omp for schedule(dynamic, 4096)
    for (int64_t i = 0; i < m; ++i) {
      const int d = (int)(rp[i + 1] - rp[i]);
      p.resize((size_t)d);
      x.resize((size_t)d);
      for (int j = 0; j < d; ++j) {
        const int64_t q = rp[i] + j;
        p[(size_t)j] = w ? mu[col[q]] * (double)w[q] : mu[col[q]];
      }
      mmq_alloc_row(p.data(), x.data(), d, (int64_t)(k ? k[i] : 1), seed, (uint64_t)(class_id ? class_id[i] : class_id_base + i), sweep);
      for (int j = 0; j < d; ++j) c[col[rp[i] + j]] += x[(size_t)j];
    }
  }
  for (int64_t t = 0; t < n; ++t) {
    int32_t v = 0;
    for (int q = 0; q < threads; ++q) v += part[(size_t)t + (size_t)n * (size_t)q];
    counts_out[t] = v;
  }
}

/* One shard's part of an EM iteration (src/mmseq.cpp:781-802 split over row blocks): acc[t] = sum over the shard's classes
 * containing t of k_i w_it / D_i (classes in ascending row order, as orc_em), returns sum_i k_i log D_i.  The caller
 * (bench.py's EM gate) sums acc and the return value over shards and applies mu' = mu acc / l. */
double orc_em_partial(int64_t m, int64_t n, const int64_t* rp, const int32_t* col, const int32_t* k, const float* w,
                      const double* mu, double* acc, int threads) {
  std::vector<double> D((size_t)m);
  double ll = 0.0;
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#endif
#pragma omp parallel for schedule(static) reduction(+ : ll) num_threads(threads)
  for (int64_t i = 0; i < m; ++i) {
    D[(size_t)i] = row_dot(rp, col, w, mu, i);
    ll += (double)(k ? k[i] : 1) * std::log(D[(size_t)i]);
  }
  for (int64_t t = 0; t < n; ++t) acc[t] = 0.0;
  for (int64_t i = 0; i < m; ++i) {
    const double ki = (double)(k ? k[i] : 1);
    for (int64_t q = rp[i]; q < rp[i + 1]; ++q) acc[col[q]] += (w ? ki * (double)w[q] : ki) / D[(size_t)i];
  }
  return ll;
}

/* Gamma step alone from given counts (used for the multi-shard replay). */
void orc_gamma_replay(int64_t n, const int32_t* counts, const double* l, double alpha, double beta,
                      uint32_t seed, uint32_t sweep, double* mu) {
  for (int64_t t = 0; t < n; ++t) {
    mmq_rng g;
    mmq_rng_init(&g, seed, MMQ_STREAM_GAMMA, (uint64_t)t, sweep);
    mu[t] = mmq_gamma(&g, alpha + (double)counts[t], beta + l[t]);
  }
}

/* n_sweeps sweeps starting at first_sweep; every sweep with sweep % stride == 0
 * stores mu into trace[t*trace_len + sweep/stride]   (src/mmseq.cpp:911-917). */
void orc_gibbs_replay(int64_t m, int64_t n, const int64_t* rp, const int32_t* col, const int32_t* k,
                      const float* w, const double* l, double alpha, double beta, uint32_t seed,
                      int64_t first_sweep, int64_t n_sweeps, int stride, int trace_len, double* mu,
                      double* trace) {
  for (int64_t s = first_sweep; s < first_sweep + n_sweeps; ++s) {
    orc_sweep_replay(m, n, rp, col, k, w, l, alpha, beta, seed, (uint32_t)s, 0, mu, nullptr, nullptr, 1);
    if (trace && s % stride == 0 && s / stride < trace_len)
      for (int64_t t = 0; t < n; ++t) trace[t * trace_len + s / stride] = mu[t];
  }
}

/* Prior draws for transcripts without hits: Gamma(alpha, 1/(beta + len*N/1e9)),
 * src/mmseq.cpp:971-978.  Stream: (seed, PRIOR, header index, slot). */
void orc_prior_replay(int64_t cnt, const int64_t* ids, const double* lscaled, double alpha, double beta,
                      uint32_t seed, int trace_len, double* out) {
  for (int64_t u = 0; u < cnt; ++u)
    for (int s = 0; s < trace_len; ++s) {
      mmq_rng g;
      mmq_rng_init(&g, seed, MMQ_STREAM_PRIOR, (uint64_t)ids[u], (uint32_t)s);
      out[u * trace_len + s] = mmq_gamma(&g, alpha, beta + lscaled[u]);
    }
}

} /* extern "C" */



extern "C" {

void orc_gsl_binomial(uint32_t seed, int64_t cnt, int64_t n, double p, int64_t* out) {
  gsl_like::Rng r(seed);
  for (int64_t i = 0; i < cnt; ++i) out[i] = gsl_like::binomial(r, p, (unsigned int)n);
}
void orc_gsl_gamma(uint32_t seed, int64_t cnt, double a, double scale, double* out) {
  gsl_like::Rng r(seed);
  for (int64_t i = 0; i < cnt; ++i) out[i] = gsl_like::gamma(r, a, scale);
}
void orc_gsl_multinomial(uint32_t seed, int64_t cnt, int d, const double* p, int64_t k, int32_t* out) {
  gsl_like::Rng r(seed);
  std::vector<unsigned int> x((size_t)d);
  for (int64_t i = 0; i < cnt; ++i) {
    gsl_like::multinomial(r, (size_t)d, (unsigned int)k, p, x.data());
    for (int j = 0; j < d; ++j) out[i * d + j] = (int32_t)x[(size_t)j];
  }
}
int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* The reference's Gibbs loop with its own data flow, src/mmseq.cpp:834-918:
 *   rg[i] = mt19937 seeded seed+i, one per thread                         (:834-838)
 *   per sweep: memset Xcolsum and the n*threads partials                   (:854-855)
 *   omp for schedule(static) over classes: p = mu[cols], gsl_ran_multinomial,
 *     store x, Xcolsums[t + n*thread] += x                                 (:864-891)
 *   barrier; Xcolsum[t] = sum over threads                                 (:893-899)
 *   mu[t] = gsl_ran_gamma(rg[thread], alpha + Xcolsum[t], 1/(beta+l[t]))   (:904-908)
 *   every stride-th sweep from 0: trace[t*L + sweep/stride] = mu[t]        (:911-917)
 * x_store (nnz ints) stands for the write-only X matrix.  Returns seconds
 * spent in the sweep loop.  threads <= 0 means omp_get_max_threads(). */
double orc_gibbs_gsl(int64_t m, int64_t n, const int64_t* rp, const int32_t* col, const int32_t* k,
                     const double* l, double alpha, double beta, int seed, int threads,
                     int64_t n_sweeps, int stride, int trace_len, double* mu, double* trace) {
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#else
  threads = 1;
#endif
  const int64_t nnz = rp[m];
  std::vector<gsl_like::Rng> rg;
  for (int i = 0; i < threads; ++i) rg.emplace_back((uint32_t)(seed + i));
  std::vector<int> Xcolsum((size_t)n), Xcolsums((size_t)n * (size_t)threads);
  std::vector<int32_t> x_store((size_t)nnz);
  int64_t maxd = 1;
  for (int64_t i = 0; i < m; ++i) maxd = std::max<int64_t>(maxd, rp[i + 1] - rp[i]);
  auto t0 = std::chrono::steady_clock::now();
  for (int64_t iter = 0; iter < n_sweeps; ++iter) {
    std::memset(Xcolsum.data(), 0, sizeof(int) * (size_t)n);
    std::memset(Xcolsums.data(), 0, sizeof(int) * (size_t)n * (size_t)threads);
#pragma omp parallel num_threads(threads)
    {
#ifdef _OPENMP
      const int tid = omp_get_thread_num();
#else
      const int tid = 0;
#endif
      std::vector<double> p((size_t)maxd);
      std::vector<unsigned int> x((size_t)maxd);
#pragma omp for schedule(static)
      for (int64_t i = 0; i < m; ++i) {
        const int64_t d = rp[i + 1] - rp[i];
        for (int64_t j = 0; j < d; ++j) p[(size_t)j] = mu[col[rp[i] + j]];
        gsl_like::multinomial(rg[(size_t)tid], (size_t)d, (unsigned int)(k ? k[i] : 1), p.data(), x.data());
        for (int64_t j = 0; j < d; ++j) {
          x_store[(size_t)(rp[i] + j)] = (int32_t)x[(size_t)j];
          Xcolsums[(size_t)col[rp[i] + j] + (size_t)n * (size_t)tid] += (int)x[(size_t)j];
        }
      }
      /* implicit barrier of the omp for */
#pragma omp for schedule(static)
      for (int64_t t = 0; t < n; ++t)
        for (int i = 0; i < threads; ++i) Xcolsum[(size_t)t] += Xcolsums[(size_t)t + (size_t)i * (size_t)n];
#pragma omp for schedule(static)
      for (int64_t t = 0; t < n; ++t)
        mu[t] = gsl_like::gamma(rg[(size_t)tid], alpha + Xcolsum[(size_t)t], 1.0 / (beta + l[t]));
    }
    if (trace && iter % stride == 0 && iter / stride < trace_len)
      for (int64_t t = 0; t < n; ++t) trace[t * trace_len + iter / stride] = mu[t];
  }
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

/* ------------------------------------------------------------------------
 * Sokal.  src/sokal.cc:33-87 restated with a plain iterative radix-2 FFT in
 * place of the reference's radix-4 routine (:96-293): FFT(x) -> power spectrum
 * -> DC bin zeroed -> FFT again (n * circular autocovariance) ->
 * var = acov0/(n(n-1)); normalise; sum = -1/3; for i: sum += rho_i - 1/6, stop
 * at the first sum < 0 with m = i+1; tau = 2(sum + (m-1)/6).  x is destroyed.
 * Return codes as the reference: 100 n > 2^21, 200 n < 4, 201 not a power of 2. */
static void fft_radix2(std::vector<double>& re, std::vector<double>& im) {
  const size_t n = re.size();
  for (size_t i = 1, j = 0; i < n; ++i) {
    size_t bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) { std::swap(re[i], re[j]); std::swap(im[i], im[j]); }
  }
  for (size_t len = 2; len <= n; len <<= 1) {
    const double ang = -2.0 * M_PI / (double)len;
    for (size_t i = 0; i < n; i += len)
      for (size_t j = 0; j < len / 2; ++j) {
        const double wr = std::cos(ang * (double)j), wi = std::sin(ang * (double)j);
        const size_t a = i + j, b = i + j + len / 2;
        const double xr = re[b] * wr - im[b] * wi, xi = re[b] * wi + im[b] * wr;
        re[b] = re[a] - xr; im[b] = im[a] - xi;
        re[a] += xr; im[a] += xi;
      }
  }
}

int orc_sokal(int n, double* x, double* var, double* tau, int* m) {
  if (n > (2 << 20)) return 100;
  if (n < 4) return 200;
  for (int t = n; t > 1; t >>= 1) if (t & 1) return 201;
  std::vector<double> re(x, x + n), im((size_t)n, 0.0);
  fft_radix2(re, im);
  for (int i = 0; i < n; ++i) { re[(size_t)i] = re[(size_t)i] * re[(size_t)i] + im[(size_t)i] * im[(size_t)i]; im[(size_t)i] = 0.0; }
  re[0] = 0.0;
  fft_radix2(re, im);
  *var = re[0] / ((double)n * (n - 1));
  const double c = 1.0 / re[0];
  for (int i = 0; i < n; ++i) x[i] = re[(size_t)i] * c;
  double sum = -0.333333333333333333333;
  *m = n + 1;
  for (int i = 0; i < n; ++i) {
    sum += x[i] - 0.166666666666666666666;
    if (sum < 0) { *m = i + 1; break; }
  }
  *tau = 2 * (sum + (*m - 1.0) / 6.0);
  return 0;
}

/* ------------------------------------------------------------------------
 * uh().  src/uh.cpp:3-26: for every set s, sum k[i] over the classes whose
 * members ALL lie in s.  Literal restatement (every set scans every class) —
 * small inputs only.  Membership is given as set_ptr/set_members (a transcript
 * may be in several sets in principle; genes and identical sets are disjoint). */
void orc_uh_literal(int64_t m, const int64_t* rp, const int32_t* col, const int32_t* k, int64_t nsets,
                    const int64_t* set_ptr, const int32_t* set_members, int32_t* out) {
  for (int64_t s = 0; s < nsets; ++s) {
    out[s] = 0;
    for (int64_t i = 0; i < m; ++i) {
      bool uniq = true;
      for (int64_t q = rp[i]; q < rp[i + 1] && uniq; ++q) {
        bool in = false;
        for (int64_t e = set_ptr[s]; e < set_ptr[s + 1]; ++e) if (set_members[e] == col[q]) { in = true; break; }
        if (!in) uniq = false;
      }
      if (uniq) out[s] += k ? k[i] : 1;
    }
  }
}


/* The product's class plan (mmseq_b200/csrc/mmq_cls_plan.h: how mmq_create re-orders a collapsed
 * shard into slots and chunks for k_alloc_cls) built on the CPU and replayed with the shared
 * sampler: the counts must equal orc_sweep_replay's.  Lets the CPU test suite check the host side
 * of the plan (slot splitting, ordering, padding, the rest / singleton lists) without a GPU.
 * stats: [applied, small classes, column slots, class slots, rest classes, rest entries, chain classes, chain column slots]. */
int orc_cls_plan_replay(int64_t m, int64_t n, const int64_t* rp, const int32_t* col, const int32_t* k, const int64_t* class_id,
                        int64_t class_id_base, const double* mu, uint32_t seed, uint32_t sweep, int32_t* counts, int64_t* stats) {
  mmq_cls_host_plan P;
  if (m == 0 || !mmq_cls_build_host(n, m, rp, col, k, class_id, class_id_base, P)) { if (stats) stats[0] = 0; return 1; }
  mmq_cls_replay_host(P, n, mu, seed, sweep, counts);
  if (stats) {
    stats[0] = 1; stats[1] = P.small_classes; stats[2] = P.packed; stats[3] = P.chunks * 32; stats[4] = P.n_rest; stats[5] = P.nnz_rest;
    stats[6] = P.n_chain; stats[7] = P.c_packed;
  }
  return 0;
}

} /* extern "C" */

/* ---- mmcollapse's trace covariance and mean-correlation scan ----------------------
 * PARITY STATUS of this part: src/mmcollapse.cpp needs Armadillo (absent from the image, no BLAS/LAPACK either), so
 * the reference's own code for it cannot be compiled here and it ships no fixtures: the two functions below restate
 *   orc_cov          get_corrs(), src/mmcollapse.cpp:514-561: R.slice(s) = cov(myM) of the 1024 x C trace matrix
 *                    (Armadillo's cov(): columns centred on their means, X^T X / (rows - 1)), non-finite entries
 *                    set to 0 (:556-558);
 *   orc_mean_corrs   mean_corrs(), src/mmcollapse.cpp:483-511: mean and sd over the samples (slices) of the
 *                    correlation of a pair, only over the samples in which both features were observed.
 * tests/test_oracle_collapse.py pins orc_cov on numpy.cov (an independent implementation of the same definition)
 * and orc_mean_corrs on a numpy transcription of :483-511. */
extern "C" {

/* M: L x C column-major (column c = the trace of feature c, as arma::mat stores it); R: C x C column-major. */
void orc_cov(const double* M, int L, int64_t C, double* R) {
  std::vector<double> X((size_t)L * (size_t)C);
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < C; ++c) {
    const double* x = M + (size_t)c * L;
    double s = 0.0;
    for (int i = 0; i < L; ++i) s += x[i];
    const double mean = s / L;
    for (int i = 0; i < L; ++i) X[(size_t)c * L + i] = x[i] - mean;
  }
  const double norm = L > 1 ? (double)(L - 1) : 1.0;
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t a = 0; a < C; ++a) {
    const double* xa = X.data() + (size_t)a * L;
    for (int64_t b = a; b < C; ++b) {
      const double* xb = X.data() + (size_t)b * L;
      double s = 0.0;
      for (int i = 0; i < L; ++i) s += xa[i] * xb[i];
      double v = s / norm;
      if (!std::isfinite(v)) v = 0.0; /* :556-558 */
      R[(size_t)a + (size_t)C * (size_t)b] = v;
      R[(size_t)b + (size_t)C * (size_t)a] = v;
    }
  }
}

/* R: C x C x ns cube (slice s at R + s*C*C), S: C x ns column-major 0/1 (observed in sample s), ts: rows to refresh.
 * V, W: C x C column-major, updated at (t, v) and (v, t) for t in ts, v in 0..C-1. */
void orc_mean_corrs(const double* R, const uint8_t* S, int64_t C, int ns, const int32_t* ts, int64_t nts, double sdpenalty,
                    double* V, double* W) {
  const size_t CC = (size_t)C * (size_t)C;
#pragma omp parallel for schedule(static)
  for (int64_t q = 0; q < nts; ++q) {
    const int64_t t = ts[q];
    for (int64_t v = 0; v < C; ++v) {
      double su = 0.0, sr = 0.0, sr2 = 0.0;
      for (int s = 0; s < ns; ++s) {
        const double* Rs = R + (size_t)s * CC;
        double r = Rs[(size_t)t + (size_t)C * (size_t)v];
        r = r / std::sqrt(Rs[(size_t)t + (size_t)C * (size_t)t]);
        r = r / std::sqrt(Rs[(size_t)v + (size_t)C * (size_t)v]);
        const double u = (double)(S[(size_t)t + (size_t)C * s] * S[(size_t)v + (size_t)C * s]);
        if (u == 0.0) r = 0.0; /* :493-497 */
        su += u; sr += u * r; sr2 += u * (r * r);
      }
      double mean = sr / su, sd = 0.0;
      if (ns > 1) {
        sd = std::sqrt((su / (su - 1.0)) * (sr2 / su - mean * mean));
        if (!std::isfinite(sd)) sd = 0.0;
      }
      mean = mean + sdpenalty * sd;
      V[(size_t)t + (size_t)C * (size_t)v] = V[(size_t)v + (size_t)C * (size_t)t] = mean;
      W[(size_t)t + (size_t)C * (size_t)v] = W[(size_t)v + (size_t)C * (size_t)t] = sd;
    }
  }
}

} /* extern "C" */
