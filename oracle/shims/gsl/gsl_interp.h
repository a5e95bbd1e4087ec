/* Oracle shim for <gsl/gsl_interp.h> / <gsl/gsl_spline.h> -- TEST INFRASTRUCTURE ONLY.
 * Natural cubic spline and linear interpolation with GSL's evaluation formulae. */
#ifndef ORACLE_GSL_INTERP_H
#define ORACLE_GSL_INTERP_H
#include <stdlib.h>
typedef struct { size_t cache, miss_count, hit_count; } gsl_interp_accel;
typedef struct { const char *name; unsigned int min_size; int kind; } gsl_interp_type;
extern const gsl_interp_type *gsl_interp_linear;
extern const gsl_interp_type *gsl_interp_cspline;
typedef struct {
    const gsl_interp_type *type;
    double *x, *y, *c;
    size_t size;
} gsl_spline;
typedef gsl_spline gsl_interp;
gsl_interp_accel *gsl_interp_accel_alloc(void);
void gsl_interp_accel_free(gsl_interp_accel *a);
gsl_spline *gsl_spline_alloc(const gsl_interp_type *T, size_t size);
int gsl_spline_init(gsl_spline *spline, const double xa[], const double ya[], size_t size);
double gsl_spline_eval(const gsl_spline *spline, double x, gsl_interp_accel *a);
void gsl_spline_free(gsl_spline *spline);
#endif
