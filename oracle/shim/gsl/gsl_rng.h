/* oracle/shim — stand-in for the GSL entry points eturro/mmseq calls (src/mmseq.cpp:836-837,
 * :880, :907, :974, :1250, :1286, :1372-1373, :1673), implemented in gsl_shim.cpp from the
 * restated algorithms of oracle/gsl_like.h.  Test infrastructure only. */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
typedef struct gsl_rng_type_s gsl_rng_type;
typedef struct gsl_rng_s gsl_rng;
extern const gsl_rng_type* gsl_rng_mt19937;
gsl_rng* gsl_rng_alloc(const gsl_rng_type* t);
void gsl_rng_set(gsl_rng* r, unsigned long seed);
void gsl_rng_free(gsl_rng* r);
#ifdef __cplusplus
}
#endif
