/* GSL RNG shim: declarations only; the oracle injects initial conditions and never
 * draws random numbers through the reference.  TEST INFRASTRUCTURE ONLY. */
#ifndef PINC_SHIM_GSL_RNG_H
#define PINC_SHIM_GSL_RNG_H
typedef struct { int dummy; } gsl_rng_type;
typedef struct { unsigned long long s; } gsl_rng;
extern const gsl_rng_type *gsl_rng_mt19937;
gsl_rng *gsl_rng_alloc(const gsl_rng_type *T);
void gsl_rng_free(gsl_rng *r);
void gsl_rng_set(const gsl_rng *r, unsigned long int seed);
double gsl_rng_uniform_pos(const gsl_rng *r);
#endif
