#ifndef ORACLE_GSL_MATH_H
#define ORACLE_GSL_MATH_H
#include <math.h>
typedef struct gsl_function_struct {
    double (*function)(double x, void *params);
    void *params;
} gsl_function;
#define GSL_FN_EVAL(F, x) (*((F)->function))(x, (F)->params)
#define GSL_DBL_EPSILON 2.2204460492503131e-16
#define GSL_DBL_MIN 2.2250738585072014e-308
#define GSL_DBL_MAX 1.7976931348623157e+308
#endif
