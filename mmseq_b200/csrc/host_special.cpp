/* host_special.cpp — the three special functions the reference takes from GSL for
 * its output tables (src/mmseq.cpp:1250/:1286 gsl_cdf_ugaussian_Pinv, :1372 gsl_sf_psi,
 * :1373 gsl_sf_psi_n(1, .)), for the host-side formatting of features WITHOUT hits
 * (closed forms, src/mmseq.cpp:1372-1395, :1523-1527).  Plain C++, no GSL. */
#include <cmath>

#include "../../include/mmq_sampler.h"

extern "C" {

/* Inverse standard normal CDF: Wichura AS241, the algorithm GSL's cdf/gaussinv.c implements. */
double mmq_host_ndtri(double p) { return mmq_ndtri(p); }

/* psi(x), x > 0: upward recurrence to x >= 10, then the asymptotic series. */
double mmq_host_digamma(double x) {
  double r = 0.0;
  while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
  const double f = 1.0 / (x * x);
  const double t = f * (-1.0 / 12.0 + f * (1.0 / 120.0 + f * (-1.0 / 252.0 + f * (1.0 / 240.0 + f * (-1.0 / 132.0 +
                   f * (691.0 / 32760.0 + f * (-1.0 / 12.0)))))));
  return r + std::log(x) - 0.5 / x + t;
}

/* psi_1(x), x > 0. */
double mmq_host_trigamma(double x) {
  double r = 0.0;
  while (x < 10.0) { r += 1.0 / (x * x); x += 1.0; }
  const double f = 1.0 / (x * x);
  const double t = 1.0 / x + 0.5 * f +
                   (1.0 / x) * f * (1.0 / 6.0 + f * (-1.0 / 30.0 + f * (1.0 / 42.0 + f * (-1.0 / 30.0 + f * (5.0 / 66.0 +
                   f * (-691.0 / 2730.0 + f * (7.0 / 6.0)))))));
  return r + t;
}

}
