/* Oracle shim for <gsl/gsl_rng.h> -- TEST INFRASTRUCTURE ONLY.
 * Restates GSL's generator interface for the five generator types the reference cycles through in
 * rng.c:55-78.  mt19937 follows Matsumoto & Nishimura's 2002 reference code exactly as GSL does
 * (seed 0 -> 4357, uniform = u32 / 2^32) and is pinned against numpy's legacy MT19937 seeding in
 * tests/test_oracle_shims.py; the other four follow their published recurrences (see gsl_shim.c). */
#ifndef ORACLE_GSL_RNG_H
#define ORACLE_GSL_RNG_H
#include <stdlib.h>
typedef struct {
    const char *name;
    unsigned long int max;
    unsigned long int min;
    size_t size;
    void (*set)(void *state, unsigned long int seed);
    unsigned long int (*get)(void *state);
    double (*get_double)(void *state);
} gsl_rng_type;
typedef struct {
    const gsl_rng_type *type;
    void *state;
} gsl_rng;
extern const gsl_rng_type *gsl_rng_mt19937;
extern const gsl_rng_type *gsl_rng_gfsr4;
extern const gsl_rng_type *gsl_rng_cmrg;
extern const gsl_rng_type *gsl_rng_mrg;
extern const gsl_rng_type *gsl_rng_taus2;
gsl_rng *gsl_rng_alloc(const gsl_rng_type *T);
void gsl_rng_free(gsl_rng *r);
void gsl_rng_set(const gsl_rng *r, unsigned long int seed);
unsigned long int gsl_rng_get(const gsl_rng *r);
double gsl_rng_uniform(const gsl_rng *r);
double gsl_rng_uniform_pos(const gsl_rng *r);
unsigned long int gsl_rng_uniform_int(const gsl_rng *r, unsigned long int n);
#endif
