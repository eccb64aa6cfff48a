/* oracle/shim/gsl_shim.cpp — the GSL stand-in linked into oracle/_ref/mmseq_ref (the unmodified
 * reference sources compiled against oracle/shim).  Samplers: oracle/gsl_like.h (MT19937 stream as
 * gsl_rng_mt19937; multinomial = conditional binomials; binomial = BTPE / inversion; gamma =
 * Marsaglia-Tsang on the polar normal).  Pinv: Wichura AS241 (what GSL's gaussinv.c implements).
 * psi / psi_1: recurrence + asymptotic series.  Test infrastructure only. */
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "../../include/mmq_sampler.h"
#include "../gsl_like.h"
#include "gsl/gsl_cdf.h"
#include "gsl/gsl_randist.h"
#include "gsl/gsl_rng.h"
#include "gsl/gsl_sf.h"

struct gsl_rng_type_s { int id; };
struct gsl_rng_s { gsl_like::Rng rng; gsl_rng_s() : rng(4357u) {} };
static const gsl_rng_type_s mt_type = {1};

extern "C" {
const gsl_rng_type* gsl_rng_mt19937 = &mt_type;
gsl_rng* gsl_rng_alloc(const gsl_rng_type*) { return new gsl_rng_s(); }
void gsl_rng_set(gsl_rng* r, unsigned long seed) { r->rng = gsl_like::Rng((uint32_t)seed); }
void gsl_rng_free(gsl_rng* r) { delete r; }
void gsl_ran_multinomial(const gsl_rng* r, size_t K, unsigned int N, const double p[], unsigned int n[]) {
  gsl_like::multinomial(const_cast<gsl_rng*>(r)->rng, K, N, p, n);
}
double gsl_ran_gamma(const gsl_rng* r, double a, double b) { return gsl_like::gamma(const_cast<gsl_rng*>(r)->rng, a, b); }
double gsl_cdf_ugaussian_Pinv(double P) { return mmq_ndtri(P); }
double gsl_sf_psi(double x) {
  double r = 0.0;
  while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
  const double f = 1.0 / (x * x);
  return r + std::log(x) - 0.5 / x +
         f * (-1.0 / 12.0 + f * (1.0 / 120.0 + f * (-1.0 / 252.0 + f * (1.0 / 240.0 + f * (-1.0 / 132.0 + f * (691.0 / 32760.0 + f * (-1.0 / 12.0)))))));
}
double gsl_sf_psi_n(int n, double x) {
  if (n != 1) { fprintf(stderr, "gsl shim: psi_n only for n == 1\n"); abort(); }
  double r = 0.0;
  while (x < 10.0) { r += 1.0 / (x * x); x += 1.0; }
  const double f = 1.0 / (x * x);
  return r + 1.0 / x + 0.5 * f +
         (1.0 / x) * f * (1.0 / 6.0 + f * (-1.0 / 30.0 + f * (1.0 / 42.0 + f * (-1.0 / 30.0 + f * (5.0 / 66.0 + f * (-691.0 / 2730.0 + f * (7.0 / 6.0)))))));
}
}
