/* Oracle shim for <gsl/gsl_sf_gamma.h> -- TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_GSL_SF_GAMMA_H
#define ORACLE_GSL_SF_GAMMA_H
double gsl_sf_gamma(const double x);
double gsl_sf_gammainv(const double x);
double gsl_sf_gamma_inc(const double a, const double x);
#endif
