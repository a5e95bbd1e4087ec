/*
 * host_recomb.cpp -- recombination-rate tables of the inhomogeneous-recombination model
 * (Sobacchi & Mesinger 2014 on the Miralda-Escude, Haehnelt & Rees 2000 density PDF).
 *
 * Replaces recombinations.c of the reference:
 *   init_MHR / free_MHR                 recombinations.c:92-138
 *   splined_recombination_rate          recombinations.c:66-90
 *   recombination_rate, MHR_rr          recombinations.c:155-215
 *   Gamma_SS                            recombinations.c:144-148 (Rahmati et al. 2013 self-shielding)
 *   A / C / beta parameter tables       recombinations.c:217-382
 *   alpha_A, alpha_B, neutral_fraction  thermochem.c:66-110
 *
 * The table is 300 redshifts (dz = 0.2) x 250 ln(Gamma12) samples of the PDF-integrated case-B rate
 * at T = 1e4 K; it is interpolated by a natural cubic spline along ln(Gamma12) only, the redshift is
 * index-sampled.  The 75 000 adaptive integrals are independent: they run over all host threads.
 * Values and spline coefficients are kept on the host (homogeneous model, one evaluation per
 * snapshot); ionize.cu uploads them (1.2 MB) for the per-cell evaluation of the inhomogeneous model.
 */
#include "host_recomb.h"

#include <cmath>
#include <cstdlib>
#include <vector>

#include "host_numerics.h"
#include "host_physics.h"
#include "rt.h"

namespace {

constexpr int A_NPTS = 60, C_NPTS = 12, BETA_NPTS = 5;
constexpr double ALPHA_B_10K = 2.59e-13; /* Constants.c:38 */

hostnum::CubicSpline g_A, g_C, g_beta;
RecombTables g_rr;
bool g_ready = false;

double C_of_z(double z) {
    if (z >= 13.0) return 1.0;
    if (z <= 2.0) return 0.558;
    return g_C.eval(z);
}
double beta_of_z(double z) {
    if (z >= 6.0) return -2.50;
    if (z <= 2.0) return -2.23;
    return g_beta.eval(z);
}
double A_of_z(double z) {
    const double hi = 2.0 + (float)A_NPTS;
    if (z >= hi) return g_A.eval(g_A.xmax()); /* the reference evaluates its spline at 62, beyond the last knot 61 */
    if (z <= 2.0) return g_A.eval(2.0);
    return g_A.eval(z);
}

double alpha_A(double T) { /* Abel et al. 1997 */
    const double t = log(T / (double)1.1604505e4);
    static const double k[10] = {-28.6130338, -0.72411256, -2.02604473e-2, -2.38086188e-3, -3.21260521e-4,
                                 -1.42150291e-5, 4.98910892e-6, 5.75561414e-7, -1.85676704e-8, -3.07113524e-9};
    double s = k[0] + k[1] * t;
    for (int i = 2; i < 10; i++) s += k[i] * pow(t, i);
    return exp(s);
}
double alpha_B(double T) { return ALPHA_B_10K * pow(T / 1.0e4, -0.75); }

/* equilibrium neutral fraction at hydrogen density n (cm^-3), T4, Gamma in 1e-12 s^-1 */
double neutral_fraction(double n, double T4, double gamma12, bool caseB) {
    const double corr_He = 1.0 / (4.0 / cosmo_params_global->Y_He - 3);
    const double alpha = caseB ? alpha_B(T4 * 1e4) : alpha_A(T4 * 1e4);
    const double gamma = gamma12 * 1e-12;
    double chi = (1 + corr_He) * n * alpha / gamma;
    if (chi < pc::TINY) return 0;
    if (chi < 1e-5) return chi;
    const double b = -2 - gamma / (n * (1 + corr_He) * alpha);
    return (-b - sqrt(b * b - 4)) / 2.0;
}

double gamma_self_shielded(double gamma_bg, double del, double T4, double z) {
    const double D_ss = 26.7 * pow(T4, 0.17) * pow((1 + z) / 10.0, -3) * pow(gamma_bg, 2.0 / 3.0);
    return gamma_bg * (0.98 * pow(1.0 + pow(del / D_ss, 1.64), -2.28) + 0.02 * pow(1.0 + del / D_ss, -0.84));
}

struct RatePars {
    double z, gamma12, T4, A, C0, beta, mean_nH;
    bool caseB;
};
/* integrand over ln(Delta): n_H P(Delta) alpha x_e^2 Delta^2, in 1e-15 s^-1 */
double rate_integrand(double lnD, const RatePars &p) {
    const double del = exp(lnD);
    const double gamma = gamma_self_shielded(p.gamma12, del, p.T4, p.z);
    const double n_H = p.mean_nH * del;
    const double x_e = 1.0 - neutral_fraction(n_H, p.T4, gamma, p.caseB);
    const double width = 2.0 * 7.61 / (3.0 * (1.0 + p.z));
    const double pdf = p.A * exp(-0.5 * pow((pow(del, -2.0 / 3.0) - p.C0) / width, 2)) * pow(del, p.beta);
    const double alpha = p.caseB ? alpha_B(p.T4 * 1e4) : alpha_A(p.T4 * 1e4);
    return 1e15 * n_H * pdf * alpha * x_e * x_e * del * del;
}

double hydrogen_density_today() { /* Constants.h:99-101 */
    const double Ho = hubble_H0();
    return 3.0 * Ho * Ho / (8.0 * M_PI * pc::G) * cosmo_params_global->OMb * (1 - cosmo_params_global->Y_He) / pc::m_p;
}

double recombination_rate(double z, double gamma12, double T4, bool caseB, int *status) {
    const RatePars p = {z, gamma12, T4, A_of_z(z), C_of_z(z), beta_of_z(z), hydrogen_density_today() * pow(1 + z, 3), caseB};
    double result = 0, err = 0;
    const int st = hostnum::qag61([&](double lnD) { return rate_integrand(lnD, p); }, log(0.01), log(200), 0, 0.01,
                                  1000, &result, &err);
    if (st != hostnum::QAG_OK) *status = st;
    return result;
}

double pdf_norm_integral(double z, int *status) {
    const double C0 = C_of_z(z), beta = beta_of_z(z);
    const double width = 2.0 * 7.61 / (3.0 * (1.0 + z));
    double result = 0, err = 0;
    const int st = hostnum::qag61(
        [&](double del) {
            const double u = pow(del, -2.0 / 3.0) - C0;
            return exp(-u * u / (2.0 * width * width)) * pow(del, beta);
        },
        1e-25, 1e25, 0, 0.001, 1000, &result, &err);
    if (st != hostnum::QAG_OK) *status = st;
    return result;
}

void build_parameter_splines(int *status) {
    static const double Cv[C_NPTS] = {0.558, 0.599, 0.611, 0.769, 0.868, 0.930, 0.964, 0.983, 0.993, 0.998, 0.999, 1.00};
    static const double Bv[BETA_NPTS] = {-2.23, -2.35, -2.48, -2.49, -2.50};
    std::vector<double> x, y;
    for (int i = 0; i < C_NPTS; i++) { x.push_back((float)i + 2.0); y.push_back(Cv[i]); }
    g_C.init(x, y);
    x.clear(); y.clear();
    for (int i = 0; i < BETA_NPTS; i++) { x.push_back((float)i + 2.0); y.push_back(Bv[i]); }
    g_beta.init(x, y);
    x.clear(); y.clear();
    for (int i = 0; i < A_NPTS; i++) {
        x.push_back(2.0 + (float)i);
        y.push_back(1.0 / pdf_norm_integral(2.0 + (float)i, status));
    }
    g_A.init(x, y);
}

}  // namespace

const RecombTables *recomb_tables() { return g_ready ? &g_rr : nullptr; }

extern "C" void init_MHR(void) {
    try {
        require_params(false);
        int status = 0;
        build_parameter_splines(&status);
        const float del_z = RECOMB_DEL_Z, del_g = RECOMB_DEL_LNGAMMA;
        for (int g = 0; g < RECOMB_NG; g++) g_rr.lnGamma[g] = RECOMB_LNGAMMA_MIN + g * del_g; /* int * float, as the reference */
        g_rr.lnGamma_max = RECOMB_LNGAMMA_MIN + del_g * (RECOMB_NG - 1);
        g_rr.y.assign((size_t)RECOMB_NZ * RECOMB_NG, 0.0);
        g_rr.c.assign((size_t)RECOMB_NZ * RECOMB_NG, 0.0);
#pragma omp parallel for schedule(dynamic, 4)
        for (int zc = 0; zc < RECOMB_NZ; zc++) {
            const float z = zc * del_z;
            int st = 0;
            std::vector<double> xs(g_rr.lnGamma, g_rr.lnGamma + RECOMB_NG), ys(RECOMB_NG);
            for (int g = 0; g < RECOMB_NG; g++) {
                const float gamma = exp(g_rr.lnGamma[g]);
                ys[g] = recombination_rate(z, gamma, 1, true, &st);
            }
            hostnum::CubicSpline sp;
            sp.init(xs, ys);
            for (int g = 0; g < RECOMB_NG; g++) {
                g_rr.y[(size_t)zc * RECOMB_NG + g] = ys[g];
                g_rr.c[(size_t)zc * RECOMB_NG + g] = sp.coeffs()[g];
            }
            if (st) {
#pragma omp atomic write
                status = st;
            }
        }
        if (status && getenv("B200_VERBOSE"))
            fprintf(stderr, "[21cmfast_b200] init_MHR: an adaptive integral returned status %d\n", status);
        g_ready = true;
    } catch (B200Error &e) {
        fprintf(stderr, "[21cmfast_b200] init_MHR: %s\n", e.msg);
    }
}

extern "C" void free_MHR(void) {
    g_ready = false;
    g_rr.y.clear();
    g_rr.c.clear();
    g_A.clear(); g_C.clear(); g_beta.clear();
}

/* recombinations.c:66-90 on the host (homogeneous model: one evaluation per snapshot) */
double recomb_rate_host(double z_eff, double gamma12_bg) {
    if (!g_ready) b200_throw(B200_TableEvaluationError, "init_MHR() has not been called");
    int z_ct = (int)(z_eff / RECOMB_DEL_Z + 0.5);
    if (z_ct < 0) z_ct = 0;
    else if (z_ct >= RECOMB_NZ) z_ct = RECOMB_NZ - 1;
    double lnG = log(gamma12_bg);
    if (lnG < RECOMB_LNGAMMA_MIN) return 0;
    if (lnG >= g_rr.lnGamma_max) lnG = g_rr.lnGamma_max - pc::FRACT_FLOAT_ERR;
    int i = (int)((lnG - RECOMB_LNGAMMA_MIN) * 10.0);
    if (i < 0) i = 0;
    if (i > RECOMB_NG - 2) i = RECOMB_NG - 2;
    while (i > 0 && g_rr.lnGamma[i] > lnG) i--;
    while (i < RECOMB_NG - 2 && g_rr.lnGamma[i + 1] <= lnG) i++;
    const double *y = &g_rr.y[(size_t)z_ct * RECOMB_NG], *c = &g_rr.c[(size_t)z_ct * RECOMB_NG];
    const double dx = g_rr.lnGamma[i + 1] - g_rr.lnGamma[i], dy = y[i + 1] - y[i], t = lnG - g_rr.lnGamma[i];
    const double b = dy / dx - dx * (c[i + 1] + 2.0 * c[i]) / 3.0;
    const double d = (c[i + 1] - c[i]) / (3.0 * dx);
    return y[i] + t * (b + t * (c[i] + t * d));
}

/* the reference's own symbol (recombinations.h:8); NaN instead of an exception across the C boundary */
extern "C" double splined_recombination_rate(double z_eff, double gamma12_bg) {
    try {
        return recomb_rate_host(z_eff, gamma12_bg);
    } catch (B200Error &) {
        return std::nan("");
    }
}
