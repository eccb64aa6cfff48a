#pragma once
#include <stddef.h>
#include "gsl_rng.h"
#ifdef __cplusplus
extern "C" {
#endif
void gsl_ran_multinomial(const gsl_rng* r, size_t K, unsigned int N, const double p[], unsigned int n[]);
double gsl_ran_gamma(const gsl_rng* r, double a, double b);
#ifdef __cplusplus
}
#endif
