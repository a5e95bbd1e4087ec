/* host_recomb.h -- recombination-rate tables (recombinations.c:35-45 of the reference), see host_recomb.cpp */
#pragma once
#include <vector>

constexpr int RECOMB_NZ = 300;               /* redshift samples, index-sampled */
constexpr int RECOMB_NG = 250;               /* ln(Gamma12) samples, splined */
constexpr float RECOMB_DEL_Z = 0.2f;         /* float in the reference: the products round like its own */
constexpr float RECOMB_DEL_LNGAMMA = 0.1f;
constexpr double RECOMB_LNGAMMA_MIN = -10.0;

struct RecombTables {
    double lnGamma[RECOMB_NG];
    double lnGamma_max;
    std::vector<double> y, c; /* [RECOMB_NZ][RECOMB_NG]: rate (1e-15 s^-1) and natural-spline c coefficients */
};

/* null until init_MHR() has run */
const RecombTables *recomb_tables();
double recomb_rate_host(double z_eff, double gamma12_bg); /* throws if init_MHR() has not run */

extern "C" {
double splined_recombination_rate(double z_eff, double gamma12_bg);
void init_MHR(void);
void free_MHR(void);
}
