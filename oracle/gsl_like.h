/* gsl_like.h — the GSL routines eturro/mmseq calls on its hot path, restated from the published
 * algorithms (GSL itself is not in this image; version unpinned in the reference, src/Makefile:16).
 * ORACLE code (test infrastructure): used by oracle/mmseq_oracle.cpp (orc_gibbs_gsl, sampler
 * cross-checks) and by oracle/shim/gsl_shim.cpp (the GSL stand-in the unmodified reference
 * sources are linked against in oracle/_ref).  Independent of include/mmq_sampler.h. */
#ifndef ORACLE_GSL_LIKE_H
#define ORACLE_GSL_LIKE_H

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <random>

namespace gsl_like {

struct Rng {
  std::mt19937 mt;
  explicit Rng(uint32_t seed) : mt(seed == 0 ? 4357u : seed) {}
  double uniform() { return mt() / 4294967296.0; }                    /* gsl_rng_uniform */
  double uniform_pos() { double x; do { x = uniform(); } while (x == 0.0); return x; }
};

static double gaussian(Rng& r) { /* polar (Box-Muller, Marsaglia) as gsl_ran_gaussian */
  double x, y, r2;
  do {
    x = -1.0 + 2.0 * r.uniform_pos();
    y = -1.0 + 2.0 * r.uniform_pos();
    r2 = x * x + y * y;
  } while (r2 > 1.0 || r2 == 0.0);
  return y * std::sqrt(-2.0 * std::log(r2) / r2);
}

static double gamma(Rng& r, double a, double b) { /* gsl_ran_gamma(r, a, b): shape a, scale b */
  if (a < 1.0) {
    double u = r.uniform_pos();
    return gamma(r, 1.0 + a, b) * std::pow(u, 1.0 / a);
  }
  double x, v, u;
  const double d = a - 1.0 / 3.0;
  const double c = (1.0 / 3.0) / std::sqrt(d);
  for (;;) {
    do {
      x = gaussian(r);
      v = 1.0 + c * x;
    } while (v <= 0.0);
    v = v * v * v;
    u = r.uniform_pos();
    if (u < 1.0 - 0.0331 * x * x * x * x) break;
    if (std::log(u) < 0.5 * x * x + d * (1.0 - v + std::log(v))) break;
  }
  return b * d * v;
}

static inline double stirling_corr(double y1) { /* BTPE's series for the log-gamma correction */
  const double y2 = y1 * y1;
  return (13860.0 - (462.0 - (132.0 - (99.0 - 140.0 / y2) / y2) / y2) / y2) / y1 / 166320.0;
}

/* Binomial(n, p): BTPE (Kachitvichyanukul & Schmeiser 1988) for n*min(p,1-p) >= 14,
 * sequential inversion below that (the split GSL's binomial_tpe.c uses). */
static unsigned int binomial(Rng& rng, double p, unsigned int n) {
  if (n == 0) return 0;
  bool flipped = false;
  if (p > 0.5) { p = 1.0 - p; flipped = true; }
  if (p <= 0.0) return flipped ? n : 0;
  const double q = 1.0 - p;
  const double s = p / q;
  const double np = n * p;
  int ix;
  if (np < 14.0) {
    const double f0 = std::pow(q, (double)n);
    for (;;) {
      double f = f0;
      double u = rng.uniform();
      for (ix = 0; ix <= 110; ++ix) {
        if (u < f) goto finish;
        u -= f;
        f *= s * (double)(n - ix) / (double)(ix + 1);
      }
    }
  } else {
    const double ffm = np + p;
    const int m = (int)ffm;
    const double xm = m + 0.5;
    const double npq = np * q;
    const double p1 = std::floor(2.195 * std::sqrt(npq) - 4.6 * q) + 0.5;
    const double xl = xm - p1;
    const double xr = xm + p1;
    const double c = 0.134 + 20.5 / (15.3 + (double)m);
    const double p2 = p1 * (1.0 + c + c);
    const double al = (ffm - xl) / (ffm - xl * p);
    const double lambda_l = al * (1.0 + 0.5 * al);
    const double ar = (xr - ffm) / (xr * q);
    const double lambda_r = ar * (1.0 + 0.5 * ar);
    const double p3 = p2 + c / lambda_l;
    const double p4 = p3 + c / lambda_r;
    double var, accept, u, v;
    for (;;) {
      u = rng.uniform() * p4;
      v = rng.uniform();
      if (u <= p1) { /* triangular region */
        ix = (int)(xm - p1 * v + u);
        goto finish;
      } else if (u <= p2) { /* parallelogram */
        const double x = xl + (u - p1) / c;
        v = v * c + 1.0 - std::fabs(x - xm) / p1;
        if (v > 1.0 || v <= 0.0) continue;
        ix = (int)x;
      } else if (u <= p3) { /* left tail */
        ix = (int)(xl + std::log(v) / lambda_l);
        if (ix < 0) continue;
        v *= ((u - p2) * lambda_l);
      } else { /* right tail */
        ix = (int)(xr - std::log(v) / lambda_r);
        if (ix > (double)n) continue;
        v *= ((u - p3) * lambda_r);
      }
      const int k = std::abs(ix - m);
      if (k <= 20) { /* explicit evaluation of f(ix)/f(m) */
        const double g = (n + 1) * s;
        double f = 1.0;
        var = v;
        if (m < ix) { for (int i = m + 1; i <= ix; ++i) f *= (g / i - s); }
        else if (m > ix) { for (int i = ix + 1; i <= m; ++i) f /= (g / i - s); }
        accept = f;
      } else { /* squeeze using upper and lower bounds on log(f(x)) */
        var = std::log(v);
        if (k < npq / 2 - 1) {
          const double amaxp = k / npq * ((k * (k / 3.0 + 0.625) + (1.0 / 6.0)) / npq + 0.5);
          const double ynorm = -(double)k * k / (2.0 * npq);
          if (var < ynorm - amaxp) goto finish;
          if (var > ynorm + amaxp) continue;
        }
        const double x1 = ix + 1.0;
        const double w = n - ix + 1.0;
        const double f1 = m + 1.0;
        const double z = n + 1.0 - m;
        accept = xm * std::log(f1 / x1) + (n - m + 0.5) * std::log(z / w) +
                 (ix - m) * std::log(w * p / (x1 * q)) + stirling_corr(f1) + stirling_corr(z) +
                 stirling_corr(x1) + stirling_corr(w);
      }
      if (var <= accept) goto finish;
    }
  }
finish:
  return flipped ? (n - (unsigned int)ix) : (unsigned int)ix;
}

/* gsl_ran_multinomial(r, K, N, p, n) */
static void multinomial(Rng& r, size_t K, unsigned int N, const double* p, unsigned int* n) {
  double norm = 0.0, sum_p = 0.0;
  unsigned int sum_n = 0;
  for (size_t k = 0; k < K; ++k) norm += p[k];
  for (size_t k = 0; k < K; ++k) {
    if (p[k] > 0.0) n[k] = binomial(r, p[k] / (norm - sum_p), N - sum_n);
    else n[k] = 0;
    sum_p += p[k];
    sum_n += n[k];
  }
}

} /* namespace gsl_like */

#endif
