/* Oracle shim for <gsl/gsl_integration.h> -- TEST INFRASTRUCTURE ONLY.
 * gsl_integration_qag = QUADPACK QAG (Piessens et al. 1983): globally adaptive bisection with a
 * Gauss-Kronrod pair selected by `key`. */
#ifndef ORACLE_GSL_INTEGRATION_H
#define ORACLE_GSL_INTEGRATION_H
#include <stdlib.h>
#include <gsl/gsl_math.h>
typedef struct {
    size_t limit, size, nrmax, i, maximum_level;
    double *alist, *blist, *rlist, *elist;
    size_t *order, *level;
} gsl_integration_workspace;
enum { GSL_INTEG_GAUSS15 = 1, GSL_INTEG_GAUSS21 = 2, GSL_INTEG_GAUSS31 = 3,
       GSL_INTEG_GAUSS41 = 4, GSL_INTEG_GAUSS51 = 5, GSL_INTEG_GAUSS61 = 6 };
gsl_integration_workspace *gsl_integration_workspace_alloc(const size_t n);
void gsl_integration_workspace_free(gsl_integration_workspace *w);
int gsl_integration_qag(const gsl_function *f, double a, double b, double epsabs, double epsrel,
                        size_t limit, int key, gsl_integration_workspace *workspace,
                        double *result, double *abserr);
#endif
