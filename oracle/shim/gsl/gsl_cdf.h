#pragma once
#ifdef __cplusplus
extern "C" {
#endif
double gsl_cdf_ugaussian_Pinv(double P);
#ifdef __cplusplus
}
#endif
