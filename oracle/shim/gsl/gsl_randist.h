#ifndef PINC_SHIM_GSL_RANDIST_H
#define PINC_SHIM_GSL_RANDIST_H
#include "gsl_rng.h"
double gsl_ran_gaussian_ziggurat(const gsl_rng *r, double sigma);
#endif
