#pragma once
#ifdef __cplusplus
extern "C" {
#endif
double gsl_sf_psi(double x);
double gsl_sf_psi_n(int n, double x);
#ifdef __cplusplus
}
#endif
