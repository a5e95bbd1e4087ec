/* Oracle shim for <gsl/gsl_randist.h> -- TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_GSL_RANDIST_H
#define ORACLE_GSL_RANDIST_H
#include <gsl/gsl_rng.h>
double gsl_ran_gaussian(const gsl_rng *r, const double sigma);
double gsl_ran_ugaussian(const gsl_rng *r);
double gsl_ran_ugaussian_tail(const gsl_rng *r, const double a);
unsigned int gsl_ran_poisson(const gsl_rng *r, double mu);
int gsl_ran_choose(const gsl_rng *r, void *dest, size_t k, void *src, size_t n, size_t size);
void gsl_ran_shuffle(const gsl_rng *r, void *base, size_t nmembm, size_t size);
#endif
