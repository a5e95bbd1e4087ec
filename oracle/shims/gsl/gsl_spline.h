#ifndef ORACLE_GSL_SPLINE_H
#define ORACLE_GSL_SPLINE_H
#include <gsl/gsl_interp.h>
#endif
